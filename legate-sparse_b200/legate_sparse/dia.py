"""dia_array — diagonal storage used to build matrices (reference legate_sparse/dia.py:65-194).

Construction-only format (host numpy); ``tocsr`` follows scipy's DIA→CSR converter — the
reference lifts the same routine (dia.py:152-190): transpose, mask of in-range NON-ZERO
entries, ``indptr = cumsum(mask.sum(axis=0))``, ``indices = row.T[mask.T]``.  For descending
offsets the reference emits unsorted rows while scipy emits sorted ones; scipy is the stated
oracle, so rows are emitted sorted (identical for the ascending offsets every test uses).
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
        assert isinstance(arg, tuple)
        data, offsets = arg
        if isinstance(offsets, int):
            offsets = numpy.full((1,), offsets)
        data, offsets = numpy.asarray(data), numpy.atleast_1d(numpy.asarray(offsets))
        if data.ndim == 1:
            data = data[None, :]
        if dtype is not None:
            data = data.astype(dtype)
        elif copy:
            data = data.copy()
        self.dtype = numpy.dtype(data.dtype)
        self.shape = tuple(int(i) for i in shape)
        self._offsets = offsets.copy() if copy else offsets
        self._data = data

    @property
    def nnz(self):
        M, N = self.shape
        nnz = 0
        for k in self.offsets:
            if k > 0:
                nnz += min(M, N - k)
            else:
                nnz += min(M + k, N)
        return int(nnz)

    @property
    def data(self):
        return self._data

    @property
    def offsets(self):
        return self._offsets

    def copy(self):
        return dia_array((self.data.copy(), self.offsets.copy()), shape=self.shape, dtype=self.dtype)

    def transpose(self, axes=None, copy=False):
        if axes is not None:
            raise ValueError(
                "Sparse matrices do not support an 'axes' parameter because swapping "
                "dimensions is the only logical permutation."
            )
        if copy:
            raise AssertionError
        num_rows, num_cols = self.shape
        max_dim = max(self.shape)
        offsets = -self.offsets
        r = numpy.arange(len(offsets), dtype=coord_ty)[:, None]
        c = numpy.arange(num_rows, dtype=coord_ty) - (offsets % max_dim)[:, None]
        pad_amount = max(0, max_dim - self.data.shape[1])
        data = numpy.hstack((self.data, numpy.zeros((self.data.shape[0], pad_amount), dtype=self.data.dtype)))
        data = data[r, c]
        return dia_array((data, offsets), shape=(num_cols, num_rows), copy=copy, dtype=self.dtype)

    T = property(transpose)

    def tocsr(self, copy=False):
        if copy:
            return self.copy().tocsr(copy=False)
        return self.transpose(copy=copy)._tocsr_transposed(copy=False)

    def _tocsr_transposed(self, copy=False):
        # self is the TRANSPOSE of the matrix being converted; the CSR of the original is the
        # CSC of self (same routine scipy's dia_matrix.tocsc runs).
        num_rows, num_cols = self.shape          # shape of the transposed operand
        out_shape = (num_cols, num_rows)
        if self.nnz == 0:
            return csr_array(out_shape, dtype=self.dtype)
        num_offsets, offset_len = self.data.shape
        offset_inds = numpy.arange(offset_len)
        row = offset_inds - self.offsets[:, None]
        mask = row >= 0
        mask &= row < num_rows
        mask &= offset_inds < num_cols
        mask &= self.data != 0
        idx_dtype = coord_ty
        indptr = numpy.zeros(num_cols + 1, dtype=idx_dtype)
        indptr[1 : offset_len + 1] = numpy.cumsum(mask.sum(axis=0, dtype=idx_dtype)[:num_cols])
        if offset_len < num_cols:
            indptr[offset_len + 1 :] = indptr[offset_len]
        indices = row.T[mask.T].astype(idx_dtype, copy=False)
        data = self.data.T[mask.T]
        out = csr_array((data, indices, indptr), shape=out_shape, dtype=self.dtype, copy=False)
        # scipy emits sorted rows; for non-ascending offsets sort inside each row
        if len(self.offsets) > 1 and numpy.any(numpy.diff(-self.offsets) < 0):
            sp = out.toscipy()
            sp.sort_indices()
            out = csr_array((sp.data, sp.indices, sp.indptr), shape=out_shape, dtype=self.dtype)
        return out


dia_matrix = dia_array
