"""Module-level exports (reference legate_sparse/module.py:40-70)."""
from .csr import csr_array  # noqa: F401
from .dia import dia_array  # noqa: F401
from .gallery import diags  # noqa: F401
from .io import mmread  # noqa: F401

# expose default types
from .types import coord_ty, nnz_ty  # noqa: F401


def is_sparse_matrix(o):
    """True for matrices created by this package (reference module.py:55-58)."""
    return any((isinstance(o, csr_array),))


issparse = is_sparse_matrix
isspmatrix = is_sparse_matrix


def isspmatrix_csr(o):
    return isinstance(o, csr_array)
