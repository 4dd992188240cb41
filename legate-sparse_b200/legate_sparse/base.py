"""CompressedBase — behaviour shared by the compressed formats
(reference legate_sparse/base.py:92-250).  The Rect<1> ``pos`` packing of the reference
(base.py:67-87, 270-296) is a Legion artifact and does not exist here: matrices store a
plain scipy-layout ``indptr``.
"""
import numpy


class CompressedBase:
    def asformat(self, format, copy=False):
        if format is None or format == getattr(self, "format", None):
            if copy:
                raise NotImplementedError
            return self
        try:
            convert_method = getattr(self, "to" + format)
        except AttributeError as e:
            raise ValueError("Format {} is unknown.".format(format)) from e
        try:
            return convert_method(copy=copy)
        except TypeError:
            return convert_method()

    def sum(self, axis=None, dtype=None, out=None):
        """Sum of the matrix elements (reference base.py:121-174: ``axis=None`` sums
        ``data``; ``axis=1/-1`` is an SpMV with a ones vector; ``axis=0`` raises)."""
        m, n = self.shape
        res_dtype = self.dtype
        if axis is None:
            return self.data.sum(dtype=res_dtype, out=out)
        if axis < 0:
            axis += 2
        if axis == 0:
            raise NotImplementedError
        ret = self @ numpy.ones((n, 1), dtype=res_dtype)
        if out is not None and out.shape != ret.shape:
            raise ValueError("dimensions do not match")
        return ret.sum(axis=axis, dtype=dtype, out=out)

    def astype(self, dtype, casting="unsafe", copy=True):
        dtype = numpy.dtype(dtype)
        if self.dtype != dtype:
            return self._with_data(self._astype_data(dtype, casting), copy=copy)
        return self.copy() if copy else self


# zero-preserving unary ufuncs applied to .data (reference base.py:210-250)
_ufuncs_with_fixed_point_at_zero = (
    "sin", "tan", "arcsin", "arctan", "sinh", "tanh", "arcsinh", "arctanh", "rint", "sign",
    "expm1", "log1p", "deg2rad", "rad2deg", "floor", "ceil", "trunc", "sqrt",
)

for _name in _ufuncs_with_fixed_point_at_zero:

    def _create_method(name):
        op = getattr(numpy, name)

        def method(self):
            return self._with_data(op(self.data))

        method.__doc__ = "Element-wise %s.\n\nSee `numpy.%s` for more information." % (name, name)
        method.__name__ = name
        return method

    setattr(CompressedBase, _name, _create_method(_name))
