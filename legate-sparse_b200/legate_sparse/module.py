"""Names exported at package level (same public names as the reference's module.py:40-70)."""
from . import csr as _csr
from .dia import dia_array  # noqa: F401
from .gallery import diags, random  # noqa: F401
from .io import mmread  # noqa: F401
from .types import coord_ty, nnz_ty  # noqa: F401

csr_array = _csr.csr_array
_SPARSE_KINDS = (_csr.csr_array,)


def is_sparse_matrix(obj) -> bool:
    """True when `obj` is a sparse matrix created by this package."""
    return isinstance(obj, _SPARSE_KINDS)


def isspmatrix_csr(obj) -> bool:
    return isinstance(obj, _csr.csr_array)


issparse = isspmatrix = is_sparse_matrix
