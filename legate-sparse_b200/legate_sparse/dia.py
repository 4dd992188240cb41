"""dia_array — diagonal storage used to assemble matrices (reference legate_sparse/dia.py:65-194).

Construction-only format, host numpy.  Layout is scipy's: ``data[d, j]`` is the entry in column j
of the diagonal ``offsets[d]``, i.e. A[j - offsets[d], j].  ``tocsr`` enumerates the stored entries
diagonal by diagonal, drops out-of-range and explicitly zero entries (the reference masks
``data != 0`` as well, dia.py:171) and orders them by (row, column) — the canonical sorted CSR that
scipy produces; for the ascending offsets used by every reference test/example this equals the
reference's own output bit for bit (tests/test_host_surface.py::test_diags_vs_reference_run).
"""
import numpy

from .base import CompressedBase
from .csr import csr_array
from .types import coord_ty


class dia_array(CompressedBase):
    format = "dia"

    def __init__(self, arg, shape=None, dtype=None, copy=False):
        if shape is None:
            raise NotImplementedError
        if not isinstance(arg, tuple):
            raise AssertionError("dia_array expects a (data, offsets) tuple")
        values, offs = arg
        offs = numpy.atleast_1d(numpy.asarray(offs if not isinstance(offs, int) else [offs]))
        values = numpy.asarray(values)
        values = values.reshape(1, -1) if values.ndim == 1 else values
        if dtype is not None:
            values = values.astype(dtype)
        elif copy:
            values = values.copy()
        self.shape = (int(shape[0]), int(shape[1]))
        self.dtype = numpy.dtype(values.dtype)
        self._values = values
        self._offs = offs.copy() if copy else offs

    # ---- stored arrays -------------------------------------------------------------------
    @property
    def data(self):
        return self._values

    @property
    def offsets(self):
        return self._offs

    @property
    def nnz(self):
        """number of positions covered by the stored diagonals (explicit zeros included)"""
        rows, cols = self.shape
        return int(sum(min(rows, cols - k) if k > 0 else min(rows + k, cols) for k in self._offs.tolist()))

    def copy(self):
        return dia_array((self._values.copy(), self._offs.copy()), shape=self.shape, dtype=self.dtype)

    # ---- transpose: B = A^T has offsets -k and data_B[d, c] = data_A[d, c + k] -----------------
    def transpose(self, axes=None, copy=False):
        if axes is not None:
            raise ValueError("Sparse matrices do not support an 'axes' parameter because swapping "
                             "dimensions is the only logical permutation.")
        if copy:
            raise AssertionError
        rows, cols = self.shape
        width = max(rows, cols)
        out = numpy.zeros((len(self._offs), rows), dtype=self.dtype)   # B has `rows` columns
        src_width = self._values.shape[1]
        for d, k in enumerate(self._offs.tolist()):
            c = numpy.arange(rows)
            src = c + k
            ok = (src >= 0) & (src < src_width) & (src < width)
            out[d, c[ok]] = self._values[d, src[ok]]
        return dia_array((out, -self._offs), shape=(cols, rows), dtype=self.dtype)

    T = property(transpose)

    # ---- DIA → CSR -----------------------------------------------------------------------
    def _tocsr_device(self):
        """DIA → CSR on the GPU (b2s_dia_to_csr_count / _fill: one thread per row walks the diagonals
        by ascending offset, drops explicit zeros — reference dia.py:159-190); the CSR arrays stay on
        the device."""
        import torch

        from . import _native as N
        from ._device import ptr, require_cuda, stream_ptr, torch_dtype, vt_enum
        from .csr import _counts_to_csr

        dev = require_cuda()
        rows, cols = self.shape
        vals = numpy.ascontiguousarray(self._values)
        ndiag, width = int(vals.shape[0]), int(vals.shape[1])
        offs = numpy.ascontiguousarray(self._offs, dtype=numpy.int64)
        order = numpy.argsort(offs, kind="stable").astype(numpy.int32)
        d_vals = torch.from_numpy(vals).to(dev)
        d_offs = torch.from_numpy(offs).to(dev)
        d_order = torch.from_numpy(order).to(dev)
        vt = vt_enum(self.dtype)
        lib = N.load()

        def count(row_nnz):
            N.check(lib.b2s_dia_to_csr_count(vt, rows, cols, ndiag, width, width, ptr(d_vals), ptr(d_offs),
                                             ptr(d_order), ptr(row_nnz), stream_ptr()), "dia_to_csr_count")

        def fill(it, indptr, idx, dat):
            N.check(lib.b2s_dia_to_csr_fill(vt, it, rows, cols, ndiag, width, width, ptr(d_vals), ptr(d_offs),
                                            ptr(d_order), ptr(indptr), ptr(idx), ptr(dat), stream_ptr()),
                    "dia_to_csr_fill")

        dat, idx, indptr = _counts_to_csr(rows, cols, count, fill, torch_dtype(self.dtype), dev)
        return csr_array((dat, idx, indptr), shape=self.shape, dtype=self.dtype)

    def tocsr(self, copy=False):
        """With a GPU the conversion runs on the device; the numpy path below serves GPU-less hosts
        (construction only — every compute entry point still requires the GPU) and is the statement
        the device kernels are tested against."""
        import torch

        from ._device import vt_enum

        rows, cols = self.shape
        if torch.cuda.is_available() and rows > 0 and cols > 0 and self._values.size > 0:
            try:
                vt_enum(self.dtype)
            except NotImplementedError:
                pass          # integer / bool matrices are representable but not computable: host path
            else:
                return self._tocsr_device()
        return self._tocsr_host()

    def _tocsr_host(self):
        rows, cols = self.shape
        r_parts, c_parts, v_parts = [], [], []
        for d, k in enumerate(self._offs.tolist()):
            j = numpy.arange(max(0, k), min(cols, rows + k, self._values.shape[1]))   # columns on this diagonal
            if j.size == 0:
                continue
            v = self._values[d, j]
            keep = v != 0
            r_parts.append((j - k)[keep]); c_parts.append(j[keep]); v_parts.append(v[keep])
        if not v_parts:
            return csr_array((rows, cols), dtype=self.dtype)
        r, c, v = numpy.concatenate(r_parts), numpy.concatenate(c_parts), numpy.concatenate(v_parts)
        order = numpy.lexsort((c, r))                      # by row, then column
        indptr = numpy.zeros(rows + 1, dtype=coord_ty)
        numpy.cumsum(numpy.bincount(r, minlength=rows), out=indptr[1:])
        return csr_array((v[order], c[order].astype(coord_ty), indptr), shape=self.shape, dtype=self.dtype)


dia_matrix = dia_array
