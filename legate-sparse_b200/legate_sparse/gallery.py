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
