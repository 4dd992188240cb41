"""Direct C-ABI calls (ctypes, raw device pointers): error codes instead of aborts, and the
small helpers on the edges of the path (reference get_diagonal.cu, pos_to_coordinates, cast and
csr_to_dense kernels)."""
import ctypes
from ctypes import byref, c_int64, c_void_p

import numpy as np
import pytest
import scipy.sparse as sp

from legate_sparse import _native as N
from oracle import oracle

pytestmark = pytest.mark.gpu


def _dev(a):
    import torch

    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def P(t):
    return c_void_p(t.data_ptr()) if t is not None else c_void_p(0)


def test_error_codes_not_aborts():
    import torch

    lib = N.load()
    S = sp.random(50, 40, density=0.2, format="csr", random_state=1)
    ip, ix, dv = _dev(S.indptr.astype(np.int64)), _dev(S.indices.astype(np.int32)), _dev(S.data)
    x, y = _dev(np.ones(40)), torch.empty(50, dtype=torch.float64, device="cuda")
    # null y
    rc = lib.b2s_spmv_csr(N.B2S_F64, N.B2S_I32, 50, 40, S.nnz, P(ip), P(ix), P(dv), P(x), c_void_p(0), c_void_p(0), 0, c_void_p(0))
    assert rc == 1 and "y is null" in N.last_error()
    # bad dtype enum, negative size, bad variant
    assert lib.b2s_spmv_csr(9, N.B2S_I32, 50, 40, S.nnz, P(ip), P(ix), P(dv), P(x), P(y), c_void_p(0), 0, c_void_p(0)) == 1
    assert lib.b2s_spmv_csr(N.B2S_F64, N.B2S_I32, -1, 40, S.nnz, P(ip), P(ix), P(dv), P(x), P(y), c_void_p(0), 0, c_void_p(0)) == 1
    assert lib.b2s_spmv_csr(N.B2S_F64, N.B2S_I32, 50, 40, S.nnz, P(ip), P(ix), P(dv), P(x), P(y), c_void_p(0), 7, c_void_p(0)) == 1
    # tile/pipe variants need a plan
    assert lib.b2s_spmv_csr(N.B2S_F64, N.B2S_I32, 50, 40, S.nnz, P(ip), P(ix), P(dv), P(x), P(y), c_void_p(0), N.B2S_SPMV_TILE, c_void_p(0)) == 1
    # plan workspace too small → B2S_ERR_WORKSPACE
    ws = torch.empty(8, dtype=torch.uint8, device="cuda")
    h = c_void_p(0)
    assert lib.b2s_spmv_plan_create(N.B2S_I32, 50, 40, S.nnz, P(ip), P(ix), P(ws), 8, c_void_p(0), byref(h)) == 3
    # plan / matrix mismatch
    nbytes = lib.b2s_spmv_plan_workspace_bytes(50, S.nnz)
    ws = torch.empty(nbytes, dtype=torch.uint8, device="cuda")
    assert lib.b2s_spmv_plan_create(N.B2S_I32, 50, 40, S.nnz, P(ip), P(ix), P(ws), nbytes, c_void_p(0), byref(h)) == 0
    assert lib.b2s_spmv_csr(N.B2S_F64, N.B2S_I32, 49, 40, S.nnz, P(ip), P(ix), P(dv), P(x), P(y), h, 0, c_void_p(0)) == 1
    assert "plan does not match" in N.last_error()
    # and the good call works (plan-free and with plan)
    assert lib.b2s_spmv_csr(N.B2S_F64, N.B2S_I32, 50, 40, S.nnz, P(ip), P(ix), P(dv), P(x), P(y), h, 0, c_void_p(0)) == 0
    assert np.allclose(y.cpu().numpy(), S @ np.ones(40), rtol=1e-13)
    assert lib.b2s_spmv_csr(N.B2S_F64, N.B2S_I32, 50, 40, S.nnz, P(ip), P(ix), P(dv), P(x), P(y), c_void_p(0), 0, c_void_p(0)) == 0
    assert np.allclose(y.cpu().numpy(), S @ np.ones(40), rtol=1e-13)
    lib.b2s_spmv_plan_destroy(h)
    # spgemm workspace too small
    c_ptr = torch.empty(51, dtype=torch.int64, device="cuda")
    a, b = c_int64(0), c_int64(0)
    assert lib.b2s_spgemm_symbolic(N.B2S_I32, 50, 40, 40, P(ip), P(ix), S.nnz, P(ip), P(ix), S.nnz, P(c_ptr), P(ws), 4,
                                   byref(a), byref(b), c_void_p(0)) == 3
    assert lib.b2s_launch_count() > 0


def test_unaligned_views_fall_back_to_tile_kernel():
    """slices that are not 16-byte aligned cannot use TMA bulk copies → AUTO picks the register-staged
    tile kernel; results are identical."""
    import torch

    lib = N.load()
    S = sp.random(3000, 2500, density=0.01, format="csr", random_state=5)
    pad = 1  # shift every array by one element → 8/4-byte aligned only
    ixs = torch.empty(S.nnz + pad, dtype=torch.int32, device="cuda")
    dvs = torch.empty(S.nnz + pad, dtype=torch.float64, device="cuda")
    ixs[pad:] = _dev(S.indices.astype(np.int32)); dvs[pad:] = _dev(S.data)
    ix, dv = ixs[pad:], dvs[pad:]
    assert ix.data_ptr() % 16 != 0
    ip = _dev(S.indptr.astype(np.int64))
    xv = np.random.default_rng(0).standard_normal(2500)
    x, y = _dev(xv), torch.empty(3000, dtype=torch.float64, device="cuda")
    nbytes = lib.b2s_spmv_plan_workspace_bytes(3000, S.nnz)
    ws = torch.empty(nbytes, dtype=torch.uint8, device="cuda")
    h = c_void_p(0)
    assert lib.b2s_spmv_plan_create(N.B2S_I32, 3000, 2500, S.nnz, P(ip), P(ix), P(ws), nbytes, c_void_p(0), byref(h)) == 0
    assert lib.b2s_spmv_csr(N.B2S_F64, N.B2S_I32, 3000, 2500, S.nnz, P(ip), P(ix), P(dv), P(x), P(y), h, 0, c_void_p(0)) == 0
    assert np.linalg.norm(y.cpu().numpy() - S @ xv) / np.linalg.norm(S @ xv) < 1e-12
    # forcing the TMA kernel on unaligned arrays is refused, not a crash
    assert lib.b2s_spmv_csr(N.B2S_F64, N.B2S_I32, 3000, 2500, S.nnz, P(ip), P(ix), P(dv), P(x), P(y), h, N.B2S_SPMV_PIPE, c_void_p(0)) == 1
    lib.b2s_spmv_plan_destroy(h)


def test_edge_helpers_vs_oracle():
    import torch

    lib = N.load()
    S = sp.random(200, 200, density=0.05, format="csr", random_state=2) + sp.eye(200, format="csr") * 3.0
    S = S.tocsr()
    ip, ix64, dv = _dev(S.indptr.astype(np.int64)), _dev(S.indices.astype(np.int64)), _dev(S.data)
    # index casts
    ix32 = torch.empty(S.nnz, dtype=torch.int32, device="cuda")
    assert lib.b2s_cast_i64_to_i32(S.nnz, P(ix64), P(ix32), c_void_p(0)) == 0
    back = torch.empty(S.nnz, dtype=torch.int64, device="cuda")
    assert lib.b2s_cast_i32_to_i64(S.nnz, P(ix32), P(back), c_void_p(0)) == 0
    assert torch.equal(back, ix64)
    # diagonal
    d = torch.empty(200, dtype=torch.float64, device="cuda")
    assert lib.b2s_csr_diagonal(N.B2S_F64, N.B2S_I32, 200, P(ip), P(ix32), P(dv), P(d), c_void_p(0)) == 0
    assert np.array_equal(d.cpu().numpy(), oracle.diagonal(S.indptr, S.indices, S.data))
    # expand rows
    rows = torch.empty(S.nnz, dtype=torch.int64, device="cuda")
    assert lib.b2s_csr_expand_rows(200, S.nnz, P(ip), P(rows), c_void_p(0)) == 0
    assert np.array_equal(rows.cpu().numpy(), oracle.expand_rows(S.indptr))
    # to dense
    out = torch.empty(200 * 200, dtype=torch.float64, device="cuda")
    assert lib.b2s_csr_to_dense(N.B2S_F64, N.B2S_I64, 200, 200, P(ip), P(ix64), P(dv), P(out), c_void_p(0)) == 0
    assert np.array_equal(out.cpu().numpy().reshape(200, 200), np.asarray(S.todense()))
