"""GPU parity: cg / gmres / cg_axpby (reference tests/integration/test_cg_solve.py,
test_gmres_solve.py, test_cg_axpby.py) + reference-run and scipy cross-checks."""
import json
import os

import numpy as np
import pytest
import scipy.sparse as sp
import scipy.sparse.linalg as spla

import legate_sparse.linalg as linalg
from legate_sparse import csr_array
from oracle import oracle
from tests import gen

pytestmark = pytest.mark.gpu
G = os.path.join(os.path.dirname(__file__), "golden")


def _spd(N=1000, seed=471014):
    import scipy.stats as stats

    class Normal(stats.rv_continuous):
        def _rvs(self, *args, size=None, random_state=None):
            return random_state.standard_normal(size)

    def sample(n, d):
        return sp.random(n, d, density=0.1, format="csr", dtype=np.float64, random_state=seed,
                         data_rvs=Normal(seed=seed)().rvs)

    A = np.asarray(sample(N, N).todense())
    A = 0.5 * (A + A.T) + N * np.eye(N)
    x = np.asarray(sample(N, 1).todense()).squeeze()
    return A, x


@pytest.mark.parametrize("y", [[2.0, 3.0]])
@pytest.mark.parametrize("x", [[0.0, 1.0]])
@pytest.mark.parametrize("isalpha", [True, False])
@pytest.mark.parametrize("negate", [True, False])
def test_cg_axpby_known_answer(y, x, isalpha, negate):
    scalar = 2.0 / 3.0
    if negate:
        scalar = -scalar
    alpha = scalar if isalpha else 1.0
    beta = 1.0 if isalpha else scalar
    expected = alpha * np.asarray(x) + beta * np.asarray(y)
    yy, xx = np.array(y), np.array(x)
    linalg.cg_axpby(yy, xx, np.array([2.0]), np.array([3.0]), isalpha=isalpha, negate=negate)
    assert np.allclose(expected, yy)
    with open(os.path.join(G, "reference_known_answers.json")) as f:
        k = json.load(f)["cg_axpby"]["expected"][f"{int(isalpha)},{int(negate)}"]
    assert np.allclose(yy, k, rtol=1e-15)
    # against the task-body restatement on a longer vector (the GPU contracts val*x+y into one
    # FMA — as nvcc does for the reference's axpby.cu — so agreement is to 1 ulp, not bitwise)
    rng = np.random.default_rng(1)
    a, b, y0, x0 = rng.random(1), rng.random(1), rng.random(1001), rng.random(1001)
    yo = oracle.axpby(y0.copy(), x0, a, b, isalpha, negate)
    yg = linalg.cg_axpby(y0.copy(), x0, a, b, isalpha=isalpha, negate=negate)
    assert np.allclose(yg, yo, rtol=4e-16, atol=1e-16)


def test_cg_solve_reference_system():
    Ad, x = _spd()
    A = csr_array(Ad)
    y = A @ x
    x_pred, iters = linalg.cg(A, y, tol=1e-8)
    assert iters > 0 and np.allclose((A @ x_pred), y, rtol=1e-8, atol=0.0)
    residuals = []
    x_cb, _ = linalg.cg(A, y, tol=1e-8, callback=lambda xk: residuals.append(y - A @ xk))
    assert np.allclose((A @ x_cb), y, rtol=1e-8, atol=0.0) and len(residuals) > 0

    def matvec(v):
        return A @ v

    x_lo, _ = linalg.cg(linalg.LinearOperator(A.shape, matvec=matvec), y, tol=1e-8)
    assert np.allclose((A @ x_lo), y, rtol=1e-8, atol=0.0)

    def matvec_out(v, out=None):
        return A.dot(v, out=out)

    x_lo2, _ = linalg.cg(linalg.LinearOperator(A.shape, matvec=matvec_out), y, tol=1e-8)
    assert np.allclose((A @ x_lo2), y, rtol=1e-8, atol=0.0)


@pytest.mark.parametrize("unfused", ["0", "1"])
def test_cg_matches_reference_run_and_oracle(monkeypatch, unfused):
    monkeypatch.setenv("LEGATE_SPARSE_CG_UNFUSED", unfused)
    z = np.load(os.path.join(G, "refrun_cg.npz"))
    n = int(z["n"])
    A = csr_array((z["A_data"], z["A_indices"], z["A_indptr"]), shape=(n, n))
    x, it = linalg.cg(A, z["b"], tol=1e-8)
    assert it == int(z["it_cg"])  # same stopping cadence as the reference loop
    assert np.linalg.norm(x - z["x_cg"]) / np.linalg.norm(z["x_cg"]) < 1e-10
    x1, it1 = linalg.cg(A, z["b"], tol=1e-8, conv_test_iters=1)
    assert it1 == int(z["it_cg1"])
    nP = int(z["nP"])
    P = csr_array((z["P_data"], z["P_indices"], z["P_indptr"]), shape=(nP, nP))
    xp, itp = linalg.cg(P, z["bp"], rtol=1e-10)
    assert itp == int(z["it_p"])
    assert np.linalg.norm(xp - z["x_p"]) / np.linalg.norm(z["x_p"]) < 1e-10


def test_cg_poisson_vs_scipy_residual():
    # north_star: CG on the 5-point Laplacian converges to the same residual as scipy within 1e-10
    N = 128
    S = gen.poisson2d_scipy(N)
    b = np.random.default_rng(2).random(N * N)
    d, o = gen.poisson2d_diagonals(N)
    import legate_sparse as sparse

    A = sparse.diags(d, o, dtype=np.float64).tocsr()
    x, iters = linalg.cg(A, b, rtol=1e-10, conv_test_iters=1)
    xs, info = spla.cg(S, b, rtol=1e-10)
    assert info == 0
    rg = np.linalg.norm(b - S @ x) / np.linalg.norm(b)
    rs = np.linalg.norm(b - S @ xs) / np.linalg.norm(b)
    assert rg <= 1e-10 and rs <= 1e-10 and abs(rg - rs) < 1e-10
    assert np.linalg.norm(x - xs) / np.linalg.norm(xs) < 1e-8
    xo, ito = oracle.cg(lambda v: S @ v, b, rtol=1e-10, conv_test_iters=1)
    assert abs(iters - ito) <= 2
    assert np.linalg.norm(x - xo) / np.linalg.norm(xo) < 1e-9


def test_cg_device_tensors_and_x0_and_preconditioner():
    import torch

    N = 64
    S = gen.poisson2d_scipy(N)
    A = csr_array(S)
    b = np.random.default_rng(4).random(N * N)
    bd = torch.from_numpy(b).cuda()
    xd, it = linalg.cg(A, bd, rtol=1e-9)
    assert isinstance(xd, torch.Tensor) and xd.is_cuda
    assert np.linalg.norm(b - S @ xd.cpu().numpy()) / np.linalg.norm(b) < 1e-8
    x0 = np.full(N * N, 0.5)
    x2, it2 = linalg.cg(A, b, x0=x0, rtol=1e-9)
    assert np.linalg.norm(b - S @ x2) / np.linalg.norm(b) < 1e-8 and np.all(x0 == 0.5)
    # Jacobi preconditioner as a user LinearOperator (numpy callables)
    dinv = 1.0 / S.diagonal()
    M = linalg.LinearOperator(S.shape, matvec=lambda v: dinv * v, dtype=np.float64)
    x3, it3 = linalg.cg(A, b, M=M, rtol=1e-9)
    assert np.linalg.norm(b - S @ x3) / np.linalg.norm(b) < 1e-8
    xo, ito = oracle.cg(lambda v: S @ v, b, M=lambda v: dinv * v, rtol=1e-9)
    assert it3 == ito
    assert np.linalg.norm(x3 - xo) / np.linalg.norm(xo) < 1e-9


@pytest.mark.parametrize("dtype,tol", [(np.float32, 2e-4), (np.complex128, 1e-9)])
def test_cg_other_dtypes(dtype, tol):
    """f32 and c128 instantiations of the fused CG kernels (the reference's dtype set)."""
    N = 48
    S = gen.poisson2d_scipy(N).astype(dtype)
    if np.dtype(dtype).kind == "c":
        S = (S + 1j * 0.0 * S).tocsr()          # real SPD stored as complex (dot is unconjugated)
    A = csr_array(S)
    assert A.dtype == np.dtype(dtype)
    b = np.random.default_rng(8).random(N * N).astype(dtype)
    x, it = linalg.cg(A, b, rtol=1e-5 if dtype == np.float32 else 1e-10)
    assert x.dtype == np.dtype(dtype) and it > 0
    assert np.linalg.norm(S @ x - b) / np.linalg.norm(b) < tol


def test_gmres_solve_reference_system():
    Ad, x = _spd()
    A = csr_array(Ad)
    y = A @ x
    x_pred, info = linalg.gmres(A, y, atol=1e-5, tol=1e-5, maxiter=300)
    assert info == 0 and np.allclose((A @ x_pred), y, atol=1e-8)


def test_gmres_matches_reference_run():
    zc = np.load(os.path.join(G, "refrun_cg.npz"))
    z = np.load(os.path.join(G, "refrun_gmres.npz"))
    n = int(zc["n"])
    A = csr_array((zc["A_data"], zc["A_indices"], zc["A_indptr"]), shape=(n, n))
    x, info = linalg.gmres(A, zc["b"], atol=1e-5, tol=1e-5, maxiter=300)
    assert info == int(z["info_g"])
    assert np.linalg.norm(x - z["x_g"]) / np.linalg.norm(z["x_g"]) < 1e-9
    nP = int(zc["nP"])
    P = csr_array((zc["P_data"], zc["P_indices"], zc["P_indptr"]), shape=(nP, nP))
    x2, info2 = linalg.gmres(P, zc["bp"], rtol=1e-8, restart=30, maxiter=3000)
    assert info2 == int(z["info_g2"])
    assert np.linalg.norm(x2 - z["x_g2"]) / np.linalg.norm(z["x_g2"]) < 1e-6
    cb = []
    linalg.gmres(P, zc["bp"], rtol=1e-8, restart=30, maxiter=3000, callback=cb.append)
    assert len(cb) > 0 and cb[-1] < 1e-7  # pr_norm callback (b_norm bug of the reference not replicated)


def test_vector_kernels_vs_numpy():
    from legate_sparse import _device as D

    rng = np.random.default_rng(6)
    for dt in (np.float32, np.float64, np.complex64, np.complex128):
        for n in (1, 7, 1000, 100003):
            x = rng.standard_normal(n).astype(dt)
            y = rng.standard_normal(n).astype(dt)
            if np.dtype(dt).kind == "c":
                x = x + 1j * rng.standard_normal(n).astype(x.real.dtype)
                y = y + 1j * rng.standard_normal(n).astype(y.real.dtype)
            xd, yd = D.to_device(x), D.to_device(y)
            tol = 1e-11 if np.dtype(dt) in (np.float64, np.complex128) else 2e-4
            assert abs(D.dot(xd, yd).cpu().numpy()[0] - x.dot(y)) <= tol * max(1.0, abs(x.dot(y)))
            assert abs(D.dot(xd, yd, conj=True).cpu().numpy()[0] - np.vdot(x, y)) <= tol * max(1.0, abs(np.vdot(x, y)))
            assert abs(D.nrm2(xd).cpu().numpy()[0] - np.linalg.norm(x)) <= tol * np.linalg.norm(x)
            # unaligned views take the scalar path
            if n > 8:
                assert abs(D.dot(xd[1:], yd[1:]).cpu().numpy()[0] - x[1:].dot(y[1:])) <= tol * max(1.0, abs(x[1:].dot(y[1:])))
