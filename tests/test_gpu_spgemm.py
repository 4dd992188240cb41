"""GPU parity: CSR x CSR SpGEMM (reference tests/integration/test_spgemm.py:25-34 + larger,
skewed and dense-accumulator cases) vs scipy.sparse and the Gustavson oracle."""
import json
import os

import numpy as np
import pytest
import scipy.sparse as sp

import legate_sparse as sparse
from oracle import oracle
from tests import gen

pytestmark = pytest.mark.gpu
G = os.path.join(os.path.dirname(__file__), "golden")


def _cmp(C, S, tol=1e-10):
    S = S.tocsr().copy()
    S.sort_indices()
    assert C.shape == S.shape
    assert np.array_equal(C.indptr, S.indptr)        # structure exact
    assert np.array_equal(C.indices, S.indices)      # sorted by column within each row
    denom = np.linalg.norm(S.data) or 1.0
    assert np.linalg.norm(C.data - S.data) / denom < tol


@pytest.mark.parametrize("N", [5, 29])
def test_csr_spgemm_reference_shapes(N):
    a, _ = gen.simple_system(N, N, seed=0)
    A = sparse.csr_array(a)
    C = A @ A.copy()
    assert np.all(np.isclose(C.todense(), a @ a))
    with pytest.raises(ValueError):
        A.dot(A, out=np.zeros(3))


def test_spgemm_known_answers():
    with open(os.path.join(G, "reference_known_answers.json")) as f:
        k = json.load(f)["readme_tridiagonal"]
    A = sparse.diags([1] * 3, [-1, 0, 1], shape=(5, 5), format="csr", dtype=np.float64)
    B = sparse.diags([3] * 3, [-1, 0, 1], shape=(5, 5), format="csr", dtype=np.float64)
    assert np.array_equal((A @ B).todense(), np.array(k["AB_dense"], dtype=np.float64))
    z = np.load(os.path.join(G, "spmv_spgemm_scipy.npz"))
    n = int(z["S_shape"][0])
    S = sparse.csr_array((z["S_data"], z["S_indices"], z["S_indptr"]), shape=(n, n))
    C = S @ S
    assert np.array_equal(C.indptr, z["C_indptr"]) and np.array_equal(C.indices, z["C_indices"])
    assert np.allclose(C.data, z["C_data"], rtol=1e-12, atol=1e-14)
    # vs the reference's CPU algorithm (first-touch order) after canonicalisation
    cp, ci, cv = oracle.spgemm(z["S_indptr"], z["S_indices"], z["S_data"], z["S_indptr"], z["S_indices"],
                               z["S_data"], n)
    O = sp.csr_array((cv, ci, cp), shape=(n, n))
    O.sort_indices()
    assert np.array_equal(C.indices, O.indices) and np.allclose(C.data, O.data, rtol=1e-12, atol=1e-14)


@pytest.mark.parametrize("dtype", [np.float32, np.float64, np.complex64, np.complex128])
@pytest.mark.parametrize("index64", ["0", "1"])
def test_spgemm_types(monkeypatch, dtype, index64):
    monkeypatch.setenv("B2S_INDEX64", index64)
    rng = np.random.default_rng(2)
    A = sp.random(300, 200, density=0.05, format="csr", random_state=1).astype(dtype)
    B = sp.random(200, 260, density=0.05, format="csr", random_state=2).astype(dtype)
    if np.dtype(dtype).kind == "c":
        A.data = A.data + 1j * rng.standard_normal(A.nnz)
        B.data = B.data + 1j * rng.standard_normal(B.nnz)
        A, B = A.astype(dtype), B.astype(dtype)
    C = sparse.csr_array(A) @ sparse.csr_array(B)
    assert C.dtype == np.dtype(dtype)
    _cmp(C, A @ B, tol=1e-10 if np.dtype(dtype) in (np.float64, np.complex128) else 1e-5)


def test_spgemm_all_row_classes():
    """rows landing in every bin: empty, <=64, <=512, <=4096 and the dense-accumulator class."""
    rng = np.random.default_rng(3)
    n = 6000
    deg = rng.integers(0, 6, size=n)
    deg[:10] = 0
    deg[100] = 300      # → hundreds of outputs (class 2/3)
    deg[200] = 2500     # → thousands of outputs (class 3/4)
    deg[300] = 5900     # → nearly dense row: class 4 (dense accumulator)
    indptr = np.zeros(n + 1, dtype=np.int64)
    np.cumsum(deg, out=indptr[1:])
    cols = np.concatenate([np.sort(rng.choice(n, size=k, replace=False)) for k in deg]).astype(np.int64)
    A = sp.csr_array((rng.standard_normal(int(indptr[-1])), cols, indptr), shape=(n, n))
    B = sp.random(n, n, density=0.002, format="csr", random_state=5)
    C = sparse.csr_array(A) @ sparse.csr_array(B)
    _cmp(C, A @ B)
    C2 = sparse.csr_array(A) @ sparse.csr_array(A)
    _cmp(C2, A @ A)
    assert C2.has_sorted_indices()


def test_spgemm_banded_and_rmat_and_rect():
    d, c, p = gen.banded_csr_arrays(4001, 11)
    S = sp.csr_array((d, c, p), shape=(4001, 4001))
    _cmp(sparse.csr_array(S) @ sparse.csr_array(S), S @ S)
    R = gen.rmat_csr(12)
    C = sparse.csr_array(R) @ sparse.csr_array(R)
    _cmp(C, R @ R)
    E = sparse.csr_array((5, 7)) @ sparse.csr_array(sp.random(7, 3, density=0.5, format="csr", random_state=1))
    assert E.nnz == 0 and E.shape == (5, 3)
    # chain R @ A @ P as examples/gmg.py does
    P = sp.random(4001, 1000, density=0.001, format="csr", random_state=3)
    Rm = P.T.tocsr()
    G1 = sparse.csr_array(Rm) @ sparse.csr_array(S) @ sparse.csr_array(P)
    _cmp(G1, Rm @ S @ P)
