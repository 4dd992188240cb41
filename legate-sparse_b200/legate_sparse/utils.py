"""dtype-promotion helpers (reference legate_sparse/utils.py:28-33,94-114,187-202).
The Legate store plumbing of the reference's utils.py is out of scope."""
import numpy

# Datatypes that spmv and spgemm operations are supported for
SUPPORTED_DATATYPES = (
    numpy.float32,
    numpy.float64,
    numpy.complex64,
    numpy.complex128,
)


def _dtype_of(a):
    import torch

    if isinstance(a, torch.Tensor):
        from ._device import np_dtype_of

        return np_dtype_of(a)
    return numpy.dtype(a.dtype)


def _size_of(a):
    import torch

    if isinstance(a, torch.Tensor):
        return a.numel()
    return a.size if hasattr(a, "size") else 2


def find_common_type(*args):
    """numpy.result_type over sparse-matrix / array dtypes, 1-element arrays treated as
    scalars (reference utils.py:94-104)."""
    from .module import is_sparse_matrix

    array_types = []
    scalar_types = []
    for array in args:
        if is_sparse_matrix(array):
            array_types.append(array.dtype)
        elif _size_of(array) == 1:
            scalar_types.append(_dtype_of(array))
        else:
            array_types.append(_dtype_of(array))
    return numpy.result_type(*array_types, *scalar_types)


def _astype(arg, dtype):
    import torch

    if isinstance(arg, torch.Tensor):
        from ._device import np_dtype_of, torch_dtype

        return arg if np_dtype_of(arg) == dtype else arg.to(torch_dtype(dtype))
    return arg.astype(dtype, copy=False)


def cast_to_common_type(*args):
    """Cast all arguments to the common dtype (no-op when already equal; utils.py:107-114)."""
    common_type = find_common_type(*args)
    return tuple(_astype(arg, common_type) for arg in args)


def is_dtype_supported(dtype) -> bool:
    """Does this datatype support SpMV and SpGEMM (reference utils.py:187-202)."""
    return numpy.dtype(dtype) in [numpy.dtype(t) for t in SUPPORTED_DATATYPES]
