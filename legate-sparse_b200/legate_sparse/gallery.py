"""diags() — assemble a sparse matrix from its diagonals.

Behaviour of the reference's gallery.diags (legate_sparse/gallery.py:77-195, a scipy.sparse.diags
derivative) including its quirks: `dtype` is mandatory, only format None / "dia" / "csr" exists,
a missing shape means a square matrix just large enough for the first diagonal, scalars broadcast
along their diagonal."""
import numpy

from .dia import dia_array

_FORMATS = (None, "dia", "csr")


def _as_diagonal_list(diagonals, offsets):
    """normalise (diagonals, offsets) to (list of 1-D arrays, 1-D int array)"""
    if numpy.isscalar(offsets):
        single = len(diagonals) == 0 or numpy.isscalar(diagonals[0])
        if not single:
            raise ValueError("Different number of diagonals and offsets.")
        return [numpy.atleast_1d(diagonals)], numpy.atleast_1d(offsets)
    bands = [numpy.atleast_1d(d) for d in diagonals]
    offs = numpy.atleast_1d(offsets)
    if len(bands) != len(offs):
        raise ValueError("Different number of diagonals and offsets.")
    return bands, offs


def diags(diagonals, offsets=0, shape=None, format=None, dtype=None):
    bands, offs = _as_diagonal_list(diagonals, offsets)
    if shape is None:
        side = len(bands[0]) + abs(int(offs[0]))
        shape = (side, side)
    if dtype is None:
        raise NotImplementedError          # the reference does not infer the dtype
    if format not in _FORMATS:
        raise NotImplementedError
    nrows, ncols = int(shape[0]), int(shape[1])
    # DIA storage: one row per diagonal, indexed by COLUMN
    width = max([0] + [min(nrows + int(k), ncols - int(k)) + max(0, int(k)) for k in offs])
    stored = numpy.zeros((len(offs), width), dtype=dtype)
    longest = min(nrows, ncols)
    for row, (band, k) in enumerate(zip(bands, (int(k) for k in offs))):
        count = min(nrows + k, ncols - k, longest)
        if count < 0:
            raise ValueError("Offset %d (index %d) out of bounds" % (k, row))
        first = max(0, k)
        if band.shape[0] not in (1, count) and band.shape[0] < count:
            raise ValueError("Diagonal length (index %d: %d at offset %d) does not agree with matrix size (%d, %d)."
                             % (row, band.shape[0], k, nrows, ncols))
        stored[row, first:first + count] = band[..., :count]
    out = dia_array((stored, offs), shape=(nrows, ncols), dtype=dtype)
    return out.tocsr() if format == "csr" else out


def _seed_of(rng) -> int:
    """64-bit seed of the counter-based device generator from scipy's `rng` argument forms."""
    if rng is None:
        return int(numpy.random.default_rng().integers(0, 2**63 - 1))
    if isinstance(rng, (int, numpy.integer)):
        return int(rng) & (2**64 - 1)
    if isinstance(rng, numpy.random.Generator):
        return int(rng.integers(0, 2**63 - 1))
    if isinstance(rng, numpy.random.RandomState):
        return int(rng.randint(0, 2**31 - 1)) * 2654435761 % (2**63)
    raise TypeError("rng must be None, an int seed, a numpy Generator or a RandomState")


def random(m, n, density=0.01, format="csr", dtype=None, rng=None, data_rvs=None, *, random_state=None,
           row_block=None):
    """Random sparse matrix with the signature of ``scipy.sparse.random`` (the reference has no
    generator: its tests densify cupynumeric random arrays, tests/integration/utils/sample.py:21-45).

    Built ON THE DEVICE by a counter-based generator (b2s_random_csr_*): exactly
    ``round(density * m * n)`` stored entries like scipy, spread over the rows as evenly as possible,
    the j-th entry of a row in the j-th of its k equal column strata (distinct, sorted columns,
    uniform over [0, n)), values uniform in [0, 1) (scipy's default ``data_rvs``).  Entry (i, j)
    depends only on (seed, i, j): every rank generates just its own row block, and the same seed
    gives the same matrix for every world size.  ``rng`` may be an int seed, a numpy Generator /
    RandomState (one 64-bit seed is drawn from it) or None (fresh entropy).  ``data_rvs(k)`` — a host
    callable as in scipy — replaces the values of this rank's k entries.  Only CSR is produced.
    ``row_block=(r0, r1)`` overrides the rows this rank generates."""
    import torch

    from . import _native as N
    from . import dist
    from ._device import ptr, require_cuda, stream_ptr, torch_dtype, vt_enum
    from .csr import _INT32_MAX, csr_array
    from .settings import settings

    if format not in (None, "csr"):
        raise NotImplementedError("Only CSR format is supported right now")
    if density < 0 or density > 1:
        raise ValueError("density expected to be 0 <= density <= 1")
    if rng is None and random_state is not None:
        rng = random_state
    m, n = int(m), int(n)
    dtype = numpy.dtype(numpy.float64 if dtype is None else dtype)
    vt = vt_enum(dtype)
    k = int(round(density * m * n))
    seed = _seed_of(rng)
    if m == 0 or n == 0 or k == 0:
        return csr_array((m, n), dtype=dtype)
    dev = require_cuda()
    lib = N.load()
    G, rank = dist.world_size(), dist.rank()
    bounds = dist.row_block_bounds(m, G)
    r0, r1 = (int(bounds[rank]), int(bounds[rank + 1])) if row_block is None else (int(row_block[0]), int(row_block[1]))
    nloc = r1 - r0
    indptr = torch.empty(nloc + 1, dtype=torch.int64, device=dev)
    N.check(lib.b2s_random_csr_rowptr(m, k, seed, r0, r1, ptr(indptr), stream_ptr()), "random_csr_rowptr")
    nnz_loc = int(lib.b2s_random_csr_block_nnz(m, k, seed, r0, r1))
    narrow = n <= _INT32_MAX and not settings.index64()
    idx = torch.empty(nnz_loc, dtype=torch.int32 if narrow else torch.int64, device=dev)
    dat = torch.empty(nnz_loc, dtype=torch_dtype(dtype), device=dev)
    N.check(lib.b2s_random_csr_fill(vt, N.B2S_I32 if narrow else N.B2S_I64, m, n, k, seed, r0, r1, 0.0, 1.0,
                                    ptr(idx), ptr(dat), stream_ptr()), "random_csr_fill")
    if data_rvs is not None:
        vals = numpy.asarray(data_rvs(nnz_loc)).astype(dtype, copy=False)
        dat = torch.from_numpy(numpy.ascontiguousarray(vals)).to(dev)
    if G == 1 and row_block is None:
        A = csr_array((dat, idx, indptr), shape=(m, n))
    else:
        A = csr_array.from_row_block(dat, idx, indptr, (m, n), row_start=r0, bounds=bounds if row_block is None else None)
    A.indices_sorted = True
    A.canonical_format = True
    return A
