"""Default index / nnz types (reference legate_sparse/types.py:20-25)."""
import numpy

coord_ty = numpy.dtype(numpy.int64)
nnz_ty = numpy.dtype(numpy.uint64)
float64 = numpy.dtype(numpy.float64)
int32 = numpy.dtype(numpy.int32)
int64 = numpy.dtype(numpy.int64)
uint64 = numpy.dtype(numpy.uint64)
