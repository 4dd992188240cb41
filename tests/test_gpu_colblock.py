"""Column-blocked SpMV operand (b2s_csr_colblock_* / b2s_spmv_colblock): the split must be a
stable permutation of the matrix and y = A_0 x; y += A_1 x; ... must equal the plain CSR SpMV of
the reference task (src/legate_sparse/array/csr/spmv.cu) — checked against the oracle and scipy,
through the raw C ABI and through csr_array."""
import ctypes
from ctypes import byref, c_int, c_int64, c_void_p

import numpy as np
import pytest
import scipy.sparse as sp

from legate_sparse import _native as N
from oracle import oracle

pytestmark = pytest.mark.gpu

VT = {np.float32: N.B2S_F32, np.float64: N.B2S_F64, np.complex64: N.B2S_C64, np.complex128: N.B2S_C128}
TOL = {np.float32: 2e-5, np.float64: 1e-12, np.complex64: 2e-5, np.complex128: 1e-12}


def _dev(a):
    import torch

    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def P(t):
    return c_void_p(t.data_ptr()) if t is not None else c_void_p(0)


def _rand_csr(rng, m, n, per_row, dtype, sort=True, empty_every=0, colrange=None):
    """rows with 0..2*per_row entries, unique columns per row; optionally unsorted / with empty rows"""
    lo, hi = colrange if colrange else (0, n)
    cnt = rng.integers(0, 2 * per_row + 1, size=m)
    if empty_every:
        cnt[::empty_every] = 0
    cnt = np.minimum(cnt, hi - lo)
    indptr = np.zeros(m + 1, dtype=np.int64)
    np.cumsum(cnt, out=indptr[1:])
    cols = np.empty(indptr[-1], dtype=np.int64)
    for r in range(m):
        c = rng.choice(hi - lo, size=cnt[r], replace=False) + lo
        cols[indptr[r]:indptr[r + 1]] = np.sort(c) if sort else c
    vals = rng.standard_normal(indptr[-1])
    if np.issubdtype(dtype, np.complexfloating):
        vals = vals + 1j * rng.standard_normal(indptr[-1])
    return indptr, cols, vals.astype(dtype)


def _colblock(lib, vt, it, m, n, ip, ix, dv, nb):
    import torch

    nbytes = lib.b2s_csr_colblock_workspace_bytes(vt, it, m, int(dv.numel()), nb)
    assert nbytes > 0
    ws = torch.empty(nbytes, dtype=torch.uint8, device="cuda")
    h = c_void_p(0)
    rc = lib.b2s_csr_colblock_create(vt, it, m, n, int(dv.numel()), P(ip), P(ix), P(dv), nb, P(ws), nbytes,
                                     c_void_p(0), byref(h))
    assert rc == 0, N.last_error()
    return h, ws


@pytest.mark.parametrize("dtype", [np.float64, np.float32, np.complex128, np.complex64])
@pytest.mark.parametrize("itype", [np.int32, np.int64])
@pytest.mark.parametrize("nb", [2, 3, 7, 32])
def test_colblock_matches_oracle(dtype, itype, nb):
    import torch

    lib = N.load()
    rng = np.random.default_rng(100 + nb)
    m, n = 3001, 4099
    indptr, cols, vals = _rand_csr(rng, m, n, 40, dtype, sort=(nb != 3), empty_every=17)
    x = rng.standard_normal(n).astype(dtype)
    if np.issubdtype(dtype, np.complexfloating):
        x = x + 1j * rng.standard_normal(n).astype(dtype)
    ref = sp.csr_matrix((vals.astype(np.complex128 if np.iscomplexobj(vals) else np.float64), cols, indptr),
                        shape=(m, n)) @ x.astype(np.complex128 if np.iscomplexobj(x) else np.float64)
    if dtype == np.float64:
        want = oracle.spmv(indptr, cols, vals, x)   # the reference task body, restated (oracle/ref_kernels.c)
        assert np.allclose(ref, want, rtol=1e-12, atol=1e-12)
    ip, ix, dv, xd = _dev(indptr), _dev(cols.astype(itype)), _dev(vals), _dev(x)
    it = N.B2S_I32 if itype == np.int32 else N.B2S_I64
    h, ws = _colblock(lib, VT[dtype], it, m, n, ip, ix, dv, nb)
    nblk, bc = c_int(0), c_int64(0)
    per = (c_int64 * 32)()
    assert lib.b2s_csr_colblock_info(h, byref(nblk), byref(bc), per) == 0
    assert nblk.value == nb and sum(per[i] for i in range(nb)) == len(vals)
    want_per = np.bincount(cols // bc.value, minlength=nb)
    assert [per[i] for i in range(nb)] == list(want_per[:nb])
    y = torch.full((m,), float("nan"), dtype=dv.dtype, device="cuda")
    assert lib.b2s_spmv_colblock(h, P(xd), P(y), c_void_p(0), c_void_p(0), c_void_p(0), 0, c_void_p(0)) == 0, N.last_error()
    got = y.cpu().numpy()
    scale = np.abs(ref).max() + 1.0
    assert np.abs(got - ref).max() <= TOL[dtype] * scale * 40
    # fused dot on the last block's launch: sum_r w[r] * y[r]
    w = _dev(rng.standard_normal(m).astype(dtype))
    dot = torch.zeros(1, dtype=dv.dtype, device="cuda")
    y.fill_(float("nan"))
    assert lib.b2s_spmv_colblock(h, P(xd), P(y), P(w), P(dot), c_void_p(0), 0, c_void_p(0)) == 0, N.last_error()
    assert np.abs(y.cpu().numpy() - ref).max() <= TOL[dtype] * scale * 40
    wd = (w.cpu().numpy().astype(ref.dtype) * ref).sum()
    assert abs(dot.cpu().numpy()[0] - wd) <= TOL[dtype] * 40 * (np.abs(ref).sum() + 1.0)
    lib.b2s_csr_colblock_destroy(h)


def test_colblock_empty_blocks_and_tiles_spanning_rows():
    """all columns inside the middle blocks (first/last blocks empty) + rows much longer than a tile"""
    import torch

    lib = N.load()
    rng = np.random.default_rng(7)
    m, n, nb = 40, 64000, 8
    indptr, cols, vals = _rand_csr(rng, m, n, 3000, np.float64, colrange=(16000, 40000))
    x = rng.standard_normal(n)
    ip, ix, dv, xd = _dev(indptr), _dev(cols.astype(np.int32)), _dev(vals), _dev(x)
    h, ws = _colblock(lib, N.B2S_F64, N.B2S_I32, m, n, ip, ix, dv, nb)
    per = (c_int64 * 32)()
    assert lib.b2s_csr_colblock_info(h, None, None, per) == 0
    assert per[0] == 0 and per[7] == 0 and per[6] == 0 and per[2] > 0
    y = torch.full((m,), float("nan"), dtype=torch.float64, device="cuda")
    assert lib.b2s_spmv_colblock(h, P(xd), P(y), c_void_p(0), c_void_p(0), c_void_p(0), 0, c_void_p(0)) == 0, N.last_error()
    want = oracle.spmv(indptr, cols, vals, x)
    assert np.allclose(y.cpu().numpy(), want, rtol=1e-11, atol=1e-11)
    lib.b2s_csr_colblock_destroy(h)


def test_colblock_errors():
    import torch

    lib = N.load()
    S = sp.random(200, 300, density=0.1, format="csr", random_state=3)
    ip, ix, dv = _dev(S.indptr.astype(np.int64)), _dev(S.indices.astype(np.int32)), _dev(S.data)
    ws = torch.empty(64, dtype=torch.uint8, device="cuda")
    h = c_void_p(0)
    assert lib.b2s_csr_colblock_create(N.B2S_F64, N.B2S_I32, 200, 300, S.nnz, P(ip), P(ix), P(dv), 1, P(ws), 64,
                                       c_void_p(0), byref(h)) == 1            # nblocks < 2
    assert lib.b2s_csr_colblock_create(N.B2S_F64, N.B2S_I32, 200, 300, S.nnz, P(ip), P(ix), P(dv), 4, P(ws), 64,
                                       c_void_p(0), byref(h)) == 3            # workspace too small
    assert lib.b2s_csr_colblock_workspace_bytes(N.B2S_F64, N.B2S_I32, 200, S.nnz, 33) == -1
    assert lib.b2s_spmv_colblock(c_void_p(0), P(dv), P(dv), c_void_p(0), c_void_p(0), c_void_p(0), 0, c_void_p(0)) == 1
    # column ids beyond ncols are reported, not scattered out of bounds
    nbytes = lib.b2s_csr_colblock_workspace_bytes(N.B2S_F64, N.B2S_I32, 200, S.nnz, 2)
    ws = torch.empty(nbytes, dtype=torch.uint8, device="cuda")
    assert lib.b2s_csr_colblock_create(N.B2S_F64, N.B2S_I32, 200, 100, S.nnz, P(ip), P(ix), P(dv), 2, P(ws), nbytes,
                                       c_void_p(0), byref(h)) == 1
    assert "outside" in N.last_error()
    # small matrices are never suggested for blocking
    nb = c_int(5)
    assert lib.b2s_csr_colblock_suggest(N.B2S_F64, N.B2S_I32, 200, 300, S.nnz, P(ip), P(ix), c_void_p(0), byref(nb)) == 0
    assert nb.value == 1


def test_csr_array_uses_colblock_when_forced(monkeypatch):
    """B2S_SPMV_COLBLOCK=N forces the blocked operand behind `A @ x`; results match scipy, and a
    later change of A.data is honoured (the blocked copy holds values)."""
    import legate_sparse as sparse

    monkeypatch.setenv("B2S_SPMV_COLBLOCK", "4")
    rng = np.random.default_rng(11)
    S = sp.random(5000, 7000, density=0.004, format="csr", random_state=9, dtype=np.float64)
    S.sort_indices()
    A = sparse.csr_array((S.data, S.indices, S.indptr), shape=S.shape)
    x = rng.standard_normal(7000)
    y = A @ x
    assert A._block().colblock not in (None, False)
    assert A._block().colblock.info()["nblocks"] == 4
    assert np.allclose(y, S @ x, rtol=1e-12, atol=1e-12)
    # pinned host tensors in/out (bench e2e path): x slices are copied while earlier blocks run
    import torch

    xh = torch.from_numpy(x).pin_memory()
    yh = torch.empty(5000, dtype=torch.float64).pin_memory()
    A.dot(xh, out=yh)
    torch.cuda.synchronize()
    assert np.allclose(yh.numpy(), S @ x, rtol=1e-12, atol=1e-12)
    # device tensors stay on the device
    yd = A @ torch.from_numpy(x).cuda()
    assert yd.is_cuda and np.allclose(yd.cpu().numpy(), S @ x, rtol=1e-12, atol=1e-12)
    A.data = S.data * 3.0
    assert np.allclose(A @ x, 3.0 * (S @ x), rtol=1e-12, atol=1e-12)
    monkeypatch.setenv("B2S_SPMV_COLBLOCK", "0")
    B = sparse.csr_array((S.data, S.indices, S.indptr), shape=S.shape)
    assert np.allclose(B @ x, S @ x, rtol=1e-12, atol=1e-12)
    assert B._block().colblock is False


def test_colblock_suggested_for_wide_random_matrix():
    """the heuristic picks blocking for rows that reach across an x larger than the slice target and
    leaves banded matrices alone (B2S_COLBLOCK_MB shrinks the target so the test stays small)."""
    import os
    import torch

    lib = N.load()
    m = n = 1 << 20
    k = 8
    rng = np.random.default_rng(5)
    cols = np.sort(rng.integers(0, n, size=(m, k)), axis=1).reshape(-1)
    indptr = np.arange(m + 1, dtype=np.int64) * k
    ip, ix = _dev(indptr), _dev(cols.astype(np.int32))
    nb = c_int(0)
    os.environ["B2S_COLBLOCK_MB"] = "2"
    try:
        assert lib.b2s_csr_colblock_suggest(N.B2S_F64, N.B2S_I32, m, n, m * k, P(ip), P(ix), c_void_p(0), byref(nb)) == 0
        assert nb.value == 4   # 8 MB of x / 2 MB slices
        band = (np.arange(m)[:, None] + np.arange(k)[None, :]).clip(0, n - 1).reshape(-1)
        ixb = _dev(band.astype(np.int32))
        assert lib.b2s_csr_colblock_suggest(N.B2S_F64, N.B2S_I32, m, n, m * k, P(ip), P(ixb), c_void_p(0), byref(nb)) == 0
        assert nb.value == 1
    finally:
        del os.environ["B2S_COLBLOCK_MB"]
    # and the blocked product of the suggested layout is right
    vals = rng.standard_normal(m * k)
    x = rng.standard_normal(n)
    dv, xd = _dev(vals), _dev(x)
    h, ws = _colblock(lib, N.B2S_F64, N.B2S_I32, m, n, ip, ix, dv, 4)
    y = torch.empty(m, dtype=torch.float64, device="cuda")
    assert lib.b2s_spmv_colblock(h, P(xd), P(y), c_void_p(0), c_void_p(0), c_void_p(0), 0, c_void_p(0)) == 0
    want = sp.csr_matrix((vals, cols, indptr), shape=(m, n)) @ x
    assert np.allclose(y.cpu().numpy(), want, rtol=1e-12, atol=1e-12)
    lib.b2s_csr_colblock_destroy(h)


@pytest.mark.parametrize("chunks", ["1", "3", "16"])
def test_host_vector_pipeline_2d_blocks(monkeypatch, chunks):
    """A.dot(x_host[, out=y_host]) with a column-blocked operand runs the 2-D pipeline
    (_device.HostPipe: early column blocks whole, the last one row chunk by row chunk, y chunks copied
    back as they finish): numpy / pinned / unpinned vectors, empty row ranges, any chunk count."""
    import torch

    import legate_sparse as sparse

    monkeypatch.setenv("B2S_SPMV_COLBLOCK", "3")
    monkeypatch.setenv("LEGATE_SPARSE_HOSTPIPE_CHUNKS", chunks)
    rng = np.random.default_rng(21)
    S = sp.random(6001, 9000, density=0.003, format="lil", random_state=4, dtype=np.float64)
    S[1000:2500, :] = 0          # a stretch of empty rows (a whole chunk at 16 chunks)
    S = S.tocsr()
    S.sort_indices()
    A = sparse.csr_array((S.data, S.indices, S.indptr), shape=S.shape)
    x = rng.standard_normal(9000)
    want = S @ x
    y = A @ x                                            # numpy in, numpy out
    assert isinstance(y, np.ndarray) and np.allclose(y, want, rtol=1e-12, atol=1e-12)
    assert A._block().hostpipe is not None
    out = np.full(6001, 7.0)
    assert A.dot(x, out=out) is out and np.allclose(out, want, rtol=1e-12, atol=1e-12)
    xh = torch.from_numpy(x).pin_memory()
    yh = torch.empty(6001, dtype=torch.float64).pin_memory()
    for s_ in (1.0, -2.0):                               # buffers of the pipeline are reused call after call
        A.dot(s_ * xh, out=yh)
        assert np.allclose(yh.numpy(), s_ * want, rtol=1e-12, atol=1e-12)
    # the device-vector path of the same matrix is untouched by the pipeline's operands
    yd = A @ torch.from_numpy(x).cuda()
    assert np.allclose(yd.cpu().numpy(), want, rtol=1e-12, atol=1e-12)
