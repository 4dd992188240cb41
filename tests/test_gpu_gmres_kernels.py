"""The tall-skinny kernels of GMRES' Arnoldi step (b2s_cgs_project / b2s_cgs_update /
b2s_vscale_inv) against numpy, through the raw C ABI — reference linalg.py:627-657:
h = V^H u ; u -= V h ; ||u|| ; v = u/||u|| ; x += V y."""
from ctypes import c_void_p

import numpy as np
import pytest

from legate_sparse import _native as N

pytestmark = pytest.mark.gpu

VT = {np.float32: N.B2S_F32, np.float64: N.B2S_F64, np.complex64: N.B2S_C64, np.complex128: N.B2S_C128}
TOL = {np.float32: 3e-5, np.float64: 1e-13, np.complex64: 3e-5, np.complex128: 1e-13}


def P(t):
    return c_void_p(t.data_ptr()) if t is not None else c_void_p(0)


def _rand(rng, shape, dtype):
    a = rng.standard_normal(shape)
    if np.issubdtype(dtype, np.complexfloating):
        a = a + 1j * rng.standard_normal(shape)
    return a.astype(dtype)


@pytest.mark.parametrize("dtype", [np.float64, np.float32, np.complex128, np.complex64])
@pytest.mark.parametrize("n,k,pad", [(1, 1, 0), (1000, 5, 0), (100003, 20, 29), (65536, 33, 0), (4099, 17, 1)])
def test_cgs_project_update_scale(dtype, n, k, pad):
    import torch

    lib = N.load()
    rng = np.random.default_rng(n + k)
    ldv = n + pad            # pad=1 / odd n: rows are not 16-byte aligned -> scalar path
    Vh = np.zeros((k, ldv), dtype=dtype)
    Vh[:, :n] = _rand(rng, (k, n), dtype) / np.sqrt(n)
    uh = _rand(rng, n, dtype)
    V, u = torch.from_numpy(Vh).cuda(), torch.from_numpy(uh).cuda()
    ws = torch.zeros(lib.b2s_cgs_workspace_bytes(), dtype=torch.uint8, device="cuda")
    h = torch.empty(k, dtype=V.dtype, device="cuda")
    rdt = torch.float32 if dtype in (np.float32, np.complex64) else torch.float64
    nrm = torch.zeros(1, dtype=rdt, device="cuda")
    wide = np.complex128 if np.issubdtype(dtype, np.complexfloating) else np.float64
    Vw, uw = Vh[:, :n].astype(wide), uh.astype(wide)

    for rep in range(2):     # second round: the workspace counter must have wrapped back to 0
        assert lib.b2s_cgs_project(VT[dtype], n, k, P(V), ldv, P(u), P(h), P(ws), c_void_p(0)) == 0, N.last_error()
        h_ref = Vw.conj() @ uw
        scale = np.abs(Vw).max() * np.abs(uw).max() * n + 1.0
        assert np.abs(h.cpu().numpy() - h_ref).max() <= TOL[dtype] * scale
    u2 = u.clone()
    assert lib.b2s_cgs_update(VT[dtype], n, k, P(V), ldv, P(h), 1, P(u2), P(nrm), P(ws), c_void_p(0)) == 0, N.last_error()
    hw = h.cpu().numpy().astype(wide)
    u_ref = uw - Vw.T @ hw
    assert np.abs(u2.cpu().numpy() - u_ref).max() <= TOL[dtype] * (np.abs(u_ref).max() + 1.0) * 50
    assert abs(float(nrm.item()) - np.linalg.norm(u_ref)) <= TOL[dtype] * 50 * (np.linalg.norm(u_ref) + 1.0)
    # x += V y (no norm, positive sign)
    x = torch.from_numpy(uh).cuda()
    assert lib.b2s_cgs_update(VT[dtype], n, k, P(V), ldv, P(h), 0, P(x), c_void_p(0), P(ws), c_void_p(0)) == 0
    x_ref = uw + Vw.T @ hw
    assert np.abs(x.cpu().numpy() - x_ref).max() <= TOL[dtype] * (np.abs(x_ref).max() + 1.0) * 50
    # v = u / ||u|| written straight into a basis row
    row = V[k - 1, :n]
    assert lib.b2s_vscale_inv(VT[dtype], n, P(u2), P(nrm), P(row), c_void_p(0)) == 0
    v_ref = u_ref / np.linalg.norm(u_ref)
    assert np.abs(row.cpu().numpy() - v_ref).max() <= TOL[dtype] * 100
    if pad:
        assert np.all(V[:, n:].cpu().numpy() == 0)   # padding untouched


def test_cgs_errors_and_k0():
    import torch

    lib = N.load()
    u = torch.ones(10, dtype=torch.float64, device="cuda")
    V = torch.ones((2, 10), dtype=torch.float64, device="cuda")
    h = torch.zeros(2, dtype=torch.float64, device="cuda")
    ws = torch.zeros(lib.b2s_cgs_workspace_bytes(), dtype=torch.uint8, device="cuda")
    assert lib.b2s_cgs_project(N.B2S_F64, 10, 2, P(V), 5, P(u), P(h), P(ws), c_void_p(0)) == 1      # ldv < n
    assert lib.b2s_cgs_project(N.B2S_F64, 10, 2, P(V), 10, P(u), c_void_p(0), P(ws), c_void_p(0)) == 1
    assert lib.b2s_cgs_update(N.B2S_F64, 10, 2000, P(V), 10, P(h), 1, P(u), c_void_p(0), P(ws), c_void_p(0)) == 1
    assert lib.b2s_cgs_project(N.B2S_F64, 10, 0, c_void_p(0), 0, P(u), c_void_p(0), c_void_p(0), c_void_p(0)) == 0
    nrm = torch.zeros(1, dtype=torch.float64, device="cuda")
    assert lib.b2s_cgs_update(N.B2S_F64, 10, 0, c_void_p(0), 0, c_void_p(0), 1, P(u), P(nrm), P(ws), c_void_p(0)) == 0
    assert abs(nrm.item() - np.sqrt(10.0)) < 1e-14      # k = 0: u unchanged, norm still computed


def test_gmres_complex_and_float32_systems():
    """gmres on the native kernels for the other value types (the reference suite only runs fp64)."""
    import scipy.sparse as sp
    import legate_sparse as sparse
    import legate_sparse.linalg as linalg

    rng = np.random.default_rng(3)
    n = 400
    for dtype, rtol in [(np.float32, 1e-4), (np.complex128, 1e-9)]:
        S = sp.random(n, n, density=0.02, format="csr", random_state=5).astype(dtype)
        if np.issubdtype(dtype, np.complexfloating):
            S = S + 1j * sp.random(n, n, density=0.02, format="csr", random_state=6)
        S = (S + sp.eye(n) * 4.0).tocsr().astype(dtype)
        b = _rand(rng, n, dtype)
        x, info = linalg.gmres(sparse.csr_array(S), b, rtol=rtol, restart=30, maxiter=600)
        assert info == 0
        assert np.linalg.norm(S @ x - b) <= rtol * np.linalg.norm(b) * 1.5
