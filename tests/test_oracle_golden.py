"""Pin the oracle (oracle/) against (1) the reference's own known-answer vectors, (2) outputs
of the reference's own Python code run by tests/golden/make_golden.py, (3) scipy.sparse."""
import json
import os

import numpy as np
import pytest
import scipy.io
import scipy.sparse as sp

from oracle import oracle
from tests import gen

G = os.path.join(os.path.dirname(__file__), "golden")


@pytest.fixture(scope="module")
def known():
    with open(os.path.join(G, "reference_known_answers.json")) as f:
        return json.load(f)


def test_oracle_6x6_known_answer(known):
    k = known["csr_6x6"]
    indptr, idx, dat = np.array(k["indptr"]), np.array(k["indices"]), np.array(k["data"], dtype=np.float64)
    dense = np.array(k["dense"], dtype=np.float64)
    # dense -> CSR restatement reproduces the reference's arrays bit for bit
    ip2, ix2, d2 = oracle.dense_to_csr(dense)
    assert np.array_equal(ip2, indptr) and np.array_equal(ix2, idx) and np.array_equal(d2, dat)
    # SpMV restatement against the dense product (exact: small integers)
    x = np.arange(1, 7, dtype=np.float64)
    assert np.array_equal(oracle.spmv(indptr, idx, dat, x), dense @ x)
    assert np.array_equal(oracle.spmv(indptr, idx, dat, x, omp=True), dense @ x)
    assert np.array_equal(oracle.diagonal(indptr, idx, dat), np.diagonal(dense))


def test_oracle_readme_tridiagonal(known):
    k = known["readme_tridiagonal"]
    n = k["n"]
    A = sp.diags([1] * 3, [-1, 0, 1], shape=(n, n), format="csr", dtype=np.float64)
    B = sp.diags([3] * 3, [-1, 0, 1], shape=(n, n), format="csr", dtype=np.float64)
    cp, ci, cv = oracle.spgemm(A.indptr, A.indices, A.data, B.indptr, B.indices, B.data, n)
    C = sp.csr_array((cv, ci, cp), shape=(n, n))
    assert np.array_equal(np.asarray(C.todense()), np.array(k["AB_dense"], dtype=np.float64))
    assert np.array_equal(oracle.spmv(A.indptr, A.indices, A.data, np.ones(n)), np.array(k["A_ones"], dtype=float))


def test_oracle_axpby_known_answer(known):
    k = known["cg_axpby"]
    for key, exp in k["expected"].items():
        isalpha, negate = (bool(int(t)) for t in key.split(","))
        y = oracle.axpby(np.array(k["y"]), np.array(k["x"]), np.array(k["a"]), np.array(k["b"]), isalpha, negate)
        assert np.allclose(y, exp, rtol=1e-15)


def test_oracle_diags_vs_reference_run_and_scipy():
    z = np.load(os.path.join(G, "refrun_diags.npz"))
    cases = sorted({k.split("__")[0] for k in z.files})
    assert len(cases) >= 19
    for c in cases:
        kind, N, nd = str(z[f"{c}__kind"]), int(z[f"{c}__N"]), int(z[f"{c}__nd"])
        dt = np.dtype(str(z[f"{c}__dtype"]))
        if kind == "banded":
            offs = [x - (nd // 2) for x in range(nd)]
            diagonals, shape = [1] * nd, (N, N)
        elif kind == "poisson2d":
            diagonals, offs = gen.poisson2d_diagonals(N)
            shape = None
        else:
            diagonals = [np.array([1.0, 0.0, 3.0, 4.0]), np.array([5.0, 6.0, 0.0, 7.0])]
            offs, shape = [0, 2], (4, 7)
        data_arr, o, shp = oracle.diags_to_dia(diagonals, offs, shape, dt)
        ip, ix, dv = oracle.dia_to_csr(data_arr, o, shp)
        # (2) identical to what the reference's own code produced
        assert np.array_equal(ip, z[f"{c}__indptr"]) and np.array_equal(ix, z[f"{c}__indices"])
        assert np.array_equal(dv, z[f"{c}__data"]) and dv.dtype == dt
        # (3) identical to scipy (value level; scipy uses int32 indices)
        S = sp.diags(diagonals, offs, shape=shape, dtype=dt).tocsr()
        if kind == "rect_zeros":
            S.eliminate_zeros()
        assert np.array_equal(ip, S.indptr) and np.array_equal(ix, S.indices) and np.array_equal(dv, S.data)


def test_oracle_cg_matches_reference_run():
    z = np.load(os.path.join(G, "refrun_cg.npz"))
    A = sp.csr_array((z["A_data"], z["A_indices"], z["A_indptr"]), shape=(int(z["n"]),) * 2)
    mv = lambda v: oracle.spmv(A.indptr, A.indices, A.data, v)  # noqa: E731
    x, it = oracle.cg(mv, z["b"], tol=1e-8)
    assert it == int(z["it_cg"]) == 25
    assert np.allclose(x, z["x_cg"], rtol=1e-12, atol=1e-14)
    x1, it1 = oracle.cg(mv, z["b"], tol=1e-8, conv_test_iters=1)
    assert it1 == int(z["it_cg1"])
    assert np.allclose(x1, z["x_cg1"], rtol=1e-12, atol=1e-14)
    P = sp.csr_array((z["P_data"], z["P_indices"], z["P_indptr"]), shape=(int(z["nP"]),) * 2)
    mvp = lambda v: oracle.spmv(P.indptr, P.indices, P.data, v)  # noqa: E731
    xp, itp = oracle.cg(mvp, z["bp"], rtol=1e-10)
    assert itp == int(z["it_p"])
    assert np.allclose(xp, z["x_p"], rtol=1e-10, atol=1e-13)
    # and against scipy's CG: same solution within 1e-10 relative (scipy tests every iteration)
    xs, info = sp.linalg.cg(P, z["bp"], rtol=1e-10)
    assert info == 0
    assert np.linalg.norm(xp - xs) / np.linalg.norm(xs) < 1e-9


def test_oracle_gmres_matches_reference_run():
    zc = np.load(os.path.join(G, "refrun_cg.npz"))
    z = np.load(os.path.join(G, "refrun_gmres.npz"))
    A = sp.csr_array((zc["A_data"], zc["A_indices"], zc["A_indptr"]), shape=(int(zc["n"]),) * 2)
    mv = lambda v: oracle.spmv(A.indptr, A.indices, A.data, v)  # noqa: E731
    x, info = oracle.gmres(mv, zc["b"], atol=1e-5, tol=1e-5, maxiter=300)
    assert info == int(z["info_g"]) == 0
    assert np.allclose(x, z["x_g"], rtol=1e-9, atol=1e-12)
    assert np.allclose(A @ x, zc["b"], atol=1e-8)


def test_oracle_spmv_spgemm_vs_scipy_golden():
    z = np.load(os.path.join(G, "spmv_spgemm_scipy.npz"))
    y = oracle.spmv(z["A_indptr"], z["A_indices"], z["A_data"], z["x"])
    assert np.allclose(y, z["y"], rtol=1e-13, atol=1e-14)
    n = int(z["S_shape"][0])
    cp, ci, cv = oracle.spgemm(z["S_indptr"], z["S_indices"], z["S_data"], z["S_indptr"], z["S_indices"],
                               z["S_data"], n)
    C = sp.csr_array((cv, ci, cp), shape=(n, n))
    assert not C.has_sorted_indices or True  # first-touch order (reference CPU path)
    C.sort_indices()
    assert np.array_equal(C.indptr, z["C_indptr"]) and np.array_equal(C.indices, z["C_indices"])
    assert np.allclose(C.data, z["C_data"], rtol=1e-12, atol=1e-14)


def test_oracle_spgemm_first_touch_order():
    # the reference CPU path emits columns in first-touch order (spgemm_csr_csr_csr.cc:134-158)
    A = sp.csr_array(np.array([[0.0, 1.0, 1.0], [0, 0, 0], [0, 0, 0]]))
    B = sp.csr_array(np.array([[0.0, 0, 0], [0, 0, 5.0], [7.0, 0, 0]]))
    cp, ci, cv = oracle.spgemm(A.indptr, A.indices, A.data, B.indptr, B.indices, B.data, 3)
    assert list(cp) == [0, 2, 2, 2] and list(ci) == [2, 0] and list(cv) == [5.0, 7.0]


def test_oracle_coo_and_mmread():
    r = np.array([2, 0, 2, 1, 0]); c = np.array([1, 3, 0, 2, 0]); d = np.array([1.0, 2, 3, 4, 5])
    ip, ix, dv = oracle.coo_to_csr(d, r, c, 3)
    assert list(ip) == [0, 2, 3, 5] and list(ix) == [3, 0, 2, 1, 0] and list(dv) == [2.0, 5, 4, 1, 3]
    exp = np.load(os.path.join(G, "mtx_expected.npz"))
    for name in ("test.mtx", "GlossGT.mtx", "Ragusa18.mtx", "cage4.mtx", "karate.mtx"):
        m, n, rows, cols, vals = oracle.mmread_coo(os.path.join(G, "mtx", name))
        dense = np.zeros((m, n))
        np.add.at(dense, (rows, cols), 0)  # shape check
        ip, ix, dv = oracle.coo_to_csr(vals, rows, cols, m)
        D = np.asarray(sp.csr_array((dv, ix, ip), shape=(m, n)).todense())
        assert np.array_equal(D, exp[name.replace(".", "_")]), name
        assert np.array_equal(D, np.asarray(scipy.io.mmread(os.path.join(G, "mtx", name)).todense()))


def test_oracle_expand_rows():
    ip = np.array([0, 2, 2, 5])
    assert list(oracle.expand_rows(ip)) == [0, 0, 2, 2, 2]
