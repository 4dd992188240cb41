"""diags() — construct a sparse matrix from diagonals
(reference legate_sparse/gallery.py:77-195, itself lifted from scipy.sparse.diags;
quirks kept: ``dtype`` is mandatory, formats limited to None/"dia"/"csr")."""
import numpy

from .dia import dia_array


def diags(diagonals, offsets=0, shape=None, format=None, dtype=None):
    # if offsets is not a sequence, assume that there's only one diagonal
    if numpy.isscalar(offsets):
        if len(diagonals) == 0 or numpy.isscalar(diagonals[0]):
            diagonals = [numpy.atleast_1d(diagonals)]
        else:
            raise ValueError("Different number of diagonals and offsets.")
    else:
        diagonals = list(map(numpy.atleast_1d, diagonals))

    offsets = numpy.atleast_1d(offsets)

    if len(diagonals) != len(offsets):
        raise ValueError("Different number of diagonals and offsets.")

    if shape is None:
        m = len(diagonals[0]) + abs(int(offsets[0]))
        shape = (m, m)

    if dtype is None:
        raise NotImplementedError

    if format is not None and format not in ["csr", "dia"]:
        raise NotImplementedError

    m, n = shape
    M = max([min(m + int(offset), n - int(offset)) + max(0, int(offset)) for offset in offsets])
    M = max(0, M)
    data_arr = numpy.zeros((len(offsets), M), dtype=dtype)
    K = min(m, n)

    for j, diagonal in enumerate(diagonals):
        offset = int(offsets[j])
        k = max(0, offset)
        length = min(m + offset, n - offset, K)
        if length < 0:
            raise ValueError("Offset %d (index %d) out of bounds" % (offset, j))
        try:
            data_arr[j, k : k + length] = diagonal[..., :length]
        except ValueError as e:
            if len(diagonal) != length and len(diagonal) != 1:
                raise ValueError(
                    "Diagonal length (index %d: %d at offset %d) does not "
                    "agree with matrix size (%d, %d)." % (j, len(diagonal), offset, m, n)
                ) from e
            raise

    dia = dia_array((data_arr, offsets), shape=(m, n), dtype=dtype)
    if format == "csr":
        return dia.tocsr()
    return dia
