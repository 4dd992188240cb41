"""Host-side logic of the legate_sparse surface (construction, dtype rules, error behaviour):
mirrors the reference's tests that need no kernel — runs without a GPU.
Reference tests: tests/integration/test_csr_from_{csr,coo,dense}.py, test_csr_to_dense.py,
test_csr_transpose.py, test_diags.py, test_unary_operation.py, test_io.py, and the
NotImplementedError cases of test_spmv.py:41-52 / test_spgemm.py:37-48."""
import json
import os

import numpy as np
import pytest
import scipy.io
import scipy.sparse as sp

import legate_sparse as sparse
import legate_sparse.linalg as linalg
from tests import gen

G = os.path.join(os.path.dirname(__file__), "golden")


@pytest.fixture(scope="module")
def known():
    with open(os.path.join(G, "reference_known_answers.json")) as f:
        return json.load(f)


def _six(known):
    k = known["csr_6x6"]
    return (np.array(k["data"], dtype=np.float64), np.array(k["indices"], dtype=np.int64),
            np.array(k["indptr"], dtype=np.int64), np.array(k["dense"], dtype=np.float64), k)


def test_csr_to_dense_known_answer(known):
    d, i, p, dense, _ = _six(known)
    A = sparse.csr_array((d, i, p), shape=(6, 6))
    assert (A.todense() == dense).all()
    assert A.nnz == 14 and A.shape == (6, 6) and A.dtype == np.float64 and A.ndim == 2 and A.dim == 2


def test_csr_from_dense_known_answer(known):
    d, i, p, dense, _ = _six(known)
    A = sparse.csr_array(dense)
    assert np.array_equal(A.data, d) and np.array_equal(A.indices, i) and np.array_equal(A.indptr, p)
    assert A.indices.dtype == np.int64 and A.indptr.dtype == np.int64


def test_unary_operations_known_answer(known):
    d, i, p, dense, k = _six(known)
    A = sparse.csr_array((d, i, p), shape=(6, 6))
    assert (np.asarray((A * 2).vals) == np.array(k["times2"], dtype=np.float64)).all()
    assert (np.asarray(A.multiply(3).vals) == np.array(k["times3"], dtype=np.float64)).all()
    assert (np.asarray((2 * A).vals) == np.array(k["times2"], dtype=np.float64)).all()
    assert np.all(np.isclose(A.todense(), A.conj().conj().todense()))
    with pytest.raises(NotImplementedError):
        A * np.ones(6)
    with pytest.raises(NotImplementedError):
        np.ones((6, 6)) @ A
    assert np.allclose(A.sqrt().todense(), np.sqrt(dense))
    assert np.allclose(A.sin().data, np.sin(d))
    assert A.sum() == d.sum()
    with pytest.raises(NotImplementedError):
        A.sum(axis=0)


@pytest.mark.parametrize("N", [7, 13])
@pytest.mark.parametrize("M", [5, 29])
def test_csr_from_coo(N, M):
    a, _ = gen.simple_system(N, M, seed=N * 100 + M)
    nz = np.argwhere(a > 0.0)
    vals = a.ravel()[a.ravel() > 0.0]
    perm = np.random.default_rng(0).permutation(len(vals))
    r, c, v = nz[perm, 0], nz[perm, 1], vals[perm]
    A = sparse.csr_array((v, (r, c)), shape=(N, M))
    assert np.all(np.isclose(a, A.todense()))
    # stable-by-row, duplicates kept, column order = input order (reference csr.py:198-219)
    from oracle import oracle

    ip, ix, dv = oracle.coo_to_csr(v, r, c, N)
    assert np.array_equal(A.indptr, ip) and np.array_equal(A.indices, ix) and np.array_equal(A.data, dv)
    with pytest.raises(AssertionError):
        sparse.csr_array((v, (r, c)))


def test_csr_from_scipy_and_empty_and_copy():
    S = sp.random(9, 11, density=0.3, format="csr", random_state=1)
    A = sparse.csr_array(S)
    assert np.array_equal(A.todense(), S.todense()) and A.indices.dtype == np.int64
    B = sparse.csr_array(A)
    B.data[:] = 0  # deep copy
    assert np.array_equal(A.todense(), S.todense())
    E = sparse.csr_array((3, 4))
    assert E.nnz == 0 and E.dtype == np.float64 and E.todense().shape == (3, 4)
    E32 = sparse.csr_array((3, 4), dtype=np.float32)
    assert E32.dtype == np.float32
    assert sparse.csr_matrix is sparse.csr_array
    assert sparse.issparse(A) and sparse.isspmatrix_csr(A) and not sparse.issparse(S)
    assert sparse.coord_ty == np.int64 and sparse.nnz_ty == np.uint64
    with pytest.raises(AttributeError):
        A.indptr = A.indptr
    assert A.tocsr() is A and A.asformat("csr") is A
    A.data = A.data * 2
    assert np.allclose(A.todense(), 2 * np.asarray(S.todense()))


@pytest.mark.parametrize("N", [5, 29])
@pytest.mark.parametrize("M", [7, 13])
@pytest.mark.parametrize("iscopy", [True, False])
def test_csr_transpose(N, M, iscopy):
    a, _ = gen.simple_system(N, M, seed=3)
    A = sparse.csr_array(a)
    assert np.all(np.isclose(a, A.T.transpose(copy=iscopy).todense()))
    assert np.array_equal(A.T.todense(), a.T)
    with pytest.raises(AssertionError):
        A.transpose(axes=(1, 0))


@pytest.mark.parametrize("N", [12, 34])
@pytest.mark.parametrize("diagonals", [3, 5])
@pytest.mark.parametrize("dtype", (np.float32, np.float64, np.complex64, np.complex128))
@pytest.mark.parametrize("fmt", ["csr", "dia"])
def test_diags(N, diagonals, dtype, fmt):
    offs = [x - (diagonals // 2) for x in range(diagonals)]
    A = sparse.diags([1] * diagonals, offs, shape=(N, N), format=fmt, dtype=dtype)
    if fmt == "dia":
        A = A.tocsr()
    B = sp.diags([1] * diagonals, offs, shape=(N, N), format="csr", dtype=dtype)
    assert np.array_equal(A.todense(), B.todense())
    # bit-exact index arrays vs scipy (value level: scipy stores int32)
    assert np.array_equal(A.indptr, B.indptr) and np.array_equal(A.indices, B.indices)
    assert np.array_equal(A.data, B.data) and A.dtype == np.dtype(dtype)


def test_diags_vs_reference_run():
    z = np.load(os.path.join(G, "refrun_diags.npz"))
    for c in sorted({k.split("__")[0] for k in z.files}):
        kind, N, nd = str(z[f"{c}__kind"]), int(z[f"{c}__N"]), int(z[f"{c}__nd"])
        dt = np.dtype(str(z[f"{c}__dtype"]))
        if kind == "banded":
            A = sparse.diags([1] * nd, [x - (nd // 2) for x in range(nd)], shape=(N, N), format="csr", dtype=dt)
        elif kind == "poisson2d":
            d, o = gen.poisson2d_diagonals(N)
            A = sparse.diags(d, o, dtype=dt).tocsr()
        else:
            A = sparse.diags([np.array([1.0, 0.0, 3.0, 4.0]), np.array([5.0, 6.0, 0.0, 7.0])], [0, 2],
                             shape=(4, 7), format="csr", dtype=dt)
        assert np.array_equal(A.indptr, z[f"{c}__indptr"]) and np.array_equal(A.indices, z[f"{c}__indices"])
        assert np.array_equal(A.data, z[f"{c}__data"]) and A.dtype == dt


def test_diags_quirks():
    with pytest.raises(NotImplementedError):
        sparse.diags([1, 2, 3], 0)  # dtype is mandatory (gallery.py:156-157)
    with pytest.raises(NotImplementedError):
        sparse.diags([1, 2, 3], 0, dtype=np.float64, format="csc")
    A = sparse.diags([1, 2, 3], 1, dtype=np.float64)  # shape inferred, DIA result
    assert A.shape == (4, 4) and isinstance(A, sparse.dia_array) and A.nnz == 3
    # descending offsets: scipy (the oracle) emits sorted rows
    B = sparse.diags([[1.0] * 4, [2.0] * 5, [3.0] * 4], [1, 0, -1], format="csr", dtype=np.float64)
    S = sp.diags([[1.0] * 4, [2.0] * 5, [3.0] * 4], [1, 0, -1], format="csr", dtype=np.float64)
    assert np.array_equal(B.indices, S.indices) and np.array_equal(B.data, S.data)


def test_poisson2d_config1_bit_exact_vs_scipy():
    # BASELINE config 1 at reduced grid (the full 1000x1000 runs in the gpu suite)
    N = 64
    d, o = gen.poisson2d_diagonals(N)
    A = sparse.diags(d, o, dtype=np.float64).tocsr()
    S = gen.poisson2d_scipy(N)
    assert A.nnz == S.nnz == 5 * N * N - 4 * N
    assert np.array_equal(A.indptr, S.indptr) and np.array_equal(A.indices, S.indices)
    assert np.array_equal(A.data, S.data)


@pytest.mark.parametrize("name", ["test.mtx", "GlossGT.mtx", "Ragusa18.mtx", "cage4.mtx", "karate.mtx"])
def test_mmread(name):
    path = os.path.join(G, "mtx", name)
    arr = sparse.mmread(path)
    s = scipy.io.mmread(path)
    assert np.array_equal(arr.todense(), np.asarray(s.todense()))
    assert arr.dtype == np.float64
    exp = np.load(os.path.join(G, "mtx_expected.npz"))[name.replace(".", "_")]
    assert np.array_equal(arr.todense(), exp)


@pytest.mark.parametrize("unsupported", ["int", "bool"])
def test_unsupported_dtypes_raise_before_anything_else(unsupported):
    d, c, p = gen.banded_csr_arrays(29, 3)
    A = sparse.csr_array((d, c, p), shape=(29, 29)).astype(unsupported)
    with pytest.raises(NotImplementedError):
        A.dot(np.ndarray((29,)))
    with pytest.raises(NotImplementedError):
        A @ A
    Af = sparse.csr_array((d, c, p), shape=(29, 29))
    with pytest.raises(NotImplementedError):
        Af.dot(np.zeros(29, dtype=np.int64))
    with pytest.raises(NotImplementedError):
        Af.dot(np.zeros((29, 3)))  # dense matrix operand is not supported


def test_no_cpu_fallback_without_gpu():
    import torch

    if torch.cuda.is_available():
        pytest.skip("GPU present")
    d, c, p = gen.banded_csr_arrays(29, 3)
    A = sparse.csr_array((d, c, p), shape=(29, 29))
    with pytest.raises(RuntimeError, match="no CUDA device|cannot load"):
        A @ np.ones(29)
    with pytest.raises(RuntimeError):
        A @ A
    with pytest.raises(RuntimeError):
        linalg.cg(A, np.ones(29))
    with pytest.raises(RuntimeError):
        linalg.cg_axpby(np.ones(2), np.ones(2), np.ones(1), np.ones(1))


def test_linear_operator_host_semantics():
    calls = []

    def mv(x):
        calls.append(x.shape)
        return 2 * x

    L = linalg.LinearOperator((4, 4), matvec=mv)
    assert type(L).__name__ == "_CustomLinearOperator" and L.dtype == np.float64
    assert np.array_equal(L.matvec(np.ones(4)), 2 * np.ones(4))
    assert L.matvec(np.ones((4, 1))).shape == (4, 1)
    out = np.zeros(4)
    L.matvec(np.arange(4.0), out=out)
    assert np.array_equal(out, 2 * np.arange(4.0))
    with pytest.raises(ValueError):
        L.matvec(np.ones(5))
    with pytest.raises(NotImplementedError):
        L.rmatvec(np.ones(4))

    def mv_out(x, out=None):
        if out is None:
            return 3 * x
        out[:] = 3 * x
        return out

    L2 = linalg.LinearOperator((4, 4), matvec=mv_out, dtype=np.float64)
    assert L2._matvec_has_out and np.array_equal(L2.matvec(np.ones(4)), 3 * np.ones(4))
    I = linalg.IdentityOperator((4, 4), dtype=np.float64)
    x = np.arange(4.0)
    y = I.matvec(x)
    assert np.array_equal(y, x) and y is not x
    assert linalg.make_linear_operator(L) is L
    with pytest.warns(RuntimeWarning):
        class Bad(linalg.LinearOperator):
            pass
        Bad(np.float64, (2, 2))
    assert linalg._get_atol_rtol(2.0, tol=1e-3, atol=0.0)[0] == pytest.approx(2e-3)
    assert linalg._get_atol_rtol(2.0, atol=1.0, rtol=1e-5)[0] == 1.0


def test_dtype_promotion_rules():
    from legate_sparse.utils import cast_to_common_type, find_common_type, is_dtype_supported

    d, c, p = gen.banded_csr_arrays(11, 3, dtype=np.float32)
    A = sparse.csr_array((d, c, p), shape=(11, 11))
    assert A.dtype == np.float32
    assert find_common_type(A, np.zeros(11, dtype=np.float64)) == np.float64
    assert find_common_type(A, np.zeros(11, dtype=np.complex64)) == np.complex64
    A2, x2 = cast_to_common_type(A, np.zeros(11, dtype=np.float64))
    assert A2.dtype == np.float64 and x2.dtype == np.float64 and A2 is not A
    A3, _ = cast_to_common_type(A, np.zeros(11, dtype=np.float32))
    assert A3 is A
    assert is_dtype_supported(np.complex128) and not is_dtype_supported(np.int32)
