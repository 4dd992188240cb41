"""Worker for tests/test_gpu_dist.py: run under torchrun with N ranks, one GPU each (NCCL)."""
import os
import sys

ROOT = os.path.normpath(os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
for p in (ROOT, os.path.join(ROOT, "legate-sparse_b200")):
    if p not in sys.path:
        sys.path.insert(0, p)

import numpy as np
import scipy.sparse as sp
import torch

import legate_sparse as sparse
import legate_sparse.linalg as linalg
from legate_sparse import dist
from oracle import oracle
from tests import gen


def relerr(a, b):
    return np.linalg.norm(np.asarray(a) - np.asarray(b)) / max(np.linalg.norm(np.asarray(b)), 1e-300)


def main():
    dist.init()
    G, rank = dist.world_size(), dist.rank()
    assert G > 1
    rng = np.random.default_rng(0)  # same on every rank (SPMD: replicated construction)
    # --- SpMV: replicated result through the public API, ragged blocks (n not divisible by G)
    n = 10007
    d, c, p = gen.random_csr_fixed(n, n, 20, seed=3)
    S = sp.csr_array((d, c, p), shape=(n, n))
    A = sparse.csr_array(S)
    x = rng.standard_normal(n)
    y = A @ x
    assert relerr(y, S @ x) < 1e-12
    blk = A._block()
    assert blk.nrows == dist.row_block_bounds(n, G)[rank + 1] - dist.row_block_bounds(n, G)[rank]
    yl = A.dot_local(torch.from_numpy(x).cuda())
    assert relerr(yl.cpu().numpy(), (S @ x)[blk.r0 : blk.r1]) < 1e-12
    # column-blocked operand + fused all-gather: the last block's launch does the peer stores
    os.environ["B2S_SPMV_COLBLOCK"] = "3"
    Ab = sparse.csr_array(S)
    yb = Ab @ x
    assert Ab._block().colblock not in (None, False)
    assert relerr(yb, S @ x) < 1e-12
    assert relerr(Ab @ (2.0 * x), 2.0 * (S @ x)) < 1e-12   # second call: buffers reused after the barrier
    # host vectors + column-blocked operand: chunked last block, y chunks copied back as they finish
    bb = Ab._block()
    yb_host = torch.empty(bb.nrows, dtype=torch.float64).pin_memory()
    for s_ in (1.0, 3.0):
        Ab.dot_local(torch.from_numpy(s_ * x).pin_memory(), out=yb_host)
        assert relerr(yb_host.numpy(), s_ * (S @ x)[bb.r0 : bb.r1]) < 1e-12
    assert bb.hostpipe is not None
    os.environ.pop("B2S_SPMV_COLBLOCK")
    # nnz-balanced partition on a power-law matrix
    d, c, p = gen.powerlaw_csr(8000, 8000, max_row=4000, seed=7)
    Sp = sp.csr_array((d, c, p), shape=(8000, 8000))
    Ap = sparse.csr_array(Sp)
    Ap.set_row_bounds(dist.nnz_balanced_bounds(p, G))
    xp = rng.standard_normal(8000)
    assert relerr(Ap @ xp, Sp @ xp) < 1e-10
    # --- CG (fused, row-sharded vectors + all-gather of p + all-reduce of the scalars)
    N = 64
    P = gen.poisson2d_scipy(N)
    Ad = sparse.csr_array(P)
    b = rng.random(N * N)
    xs, it = linalg.cg(Ad, b, rtol=1e-10)
    xo, ito = oracle.cg(lambda v: P @ v, b, rtol=1e-10)
    assert it == ito, (it, ito)
    assert relerr(xs, xo) < 1e-9
    # unfused path (replicated vectors)
    os.environ["LEGATE_SPARSE_CG_UNFUSED"] = "1"
    xu, itu = linalg.cg(Ad, b, rtol=1e-10)
    os.environ["LEGATE_SPARSE_CG_UNFUSED"] = "0"
    assert itu == ito and relerr(xu, xo) < 1e-9
    # halo exchange with a WIDE band (image reaches several hundred rows into the neighbour block)
    # and ragged nnz-balanced blocks
    nb = 3001
    half = 350
    W = sp.diags([np.full(nb - abs(o), -1.0 / (1 + abs(o))) for o in range(-half, half + 1, 7)],
                 list(range(-half, half + 1, 7)), format="csr", dtype=np.float64)
    W = (W + sp.eye(nb, format="csr") * (abs(W).sum(axis=1).max() + 1.0)).tocsr()
    Aw = sparse.csr_array(W)
    Aw.set_row_bounds(dist.nnz_balanced_bounds(W.indptr.astype(np.int64), G))
    bw = rng.random(nb)
    xw, itw = linalg.cg(Aw, bw, rtol=1e-11)
    xow, itow = oracle.cg(lambda v: W @ v, bw, rtol=1e-11)
    assert itw == itow, (itw, itow)
    assert relerr(xw, xow) < 1e-10
    assert relerr(W @ xw, bw) < 1e-10
    # --- SpGEMM: row blocks of A x replicated B, C all-gathered(v)
    R = gen.rmat_csr(10)
    Rd = sparse.csr_array(R)
    C = Rd @ Rd
    E = (R @ R).tocsr()
    E.sort_indices()
    # C stays ROW-SHARDED (reference spgemm_csr_csr_csr.cu:43-62,317-332): only per-rank nnz travelled
    assert C._g_data is None and C._h_data is None and C._blk is not None
    assert C.nnz == E.nnz
    b0, b1 = int(C.row_bounds()[rank]), int(C.row_bounds()[rank + 1])
    assert C.nnz_offset() == int(E.indptr[b0]) and C._blk.nnz == int(E.indptr[b1] - E.indptr[b0])
    # row-sharded C as the A of the next product (no gather), then as its B (gathered on request)
    C2 = C @ Rd
    E2 = (E @ R).tocsr(); E2.sort_indices()
    assert C2.nnz == E2.nnz and C._g_data is None
    C3 = Rd @ C
    E3 = (R @ E).tocsr(); E3.sort_indices()
    assert np.array_equal(C3.indptr, E3.indptr) and np.array_equal(C3.indices, E3.indices)
    assert relerr(C3.data, E3.data) < 1e-12
    yC = C @ np.ones(R.shape[0])                      # SpMV with the row-sharded product
    assert relerr(yC, E @ np.ones(R.shape[0])) < 1e-12
    assert np.array_equal(C.indptr, E.indptr) and np.array_equal(C.indices, E.indices)
    assert relerr(C.data, E.data) < 1e-12
    assert np.array_equal(C2.indptr, E2.indptr) and relerr(C2.data, E2.data) < 1e-12
    # distributed diagonal
    assert np.allclose(Ad.diagonal(), P.diagonal())
    # --- host vectors: x uploaded in 1/G slices + NVLink all-gather; dot_local returns this rank's rows
    yh = A.dot_local(x)
    assert relerr(yh, (S @ x)[blk.r0 : blk.r1]) < 1e-12
    out_blk = torch.empty(blk.nrows, dtype=torch.float64).pin_memory()
    A.dot_local(torch.from_numpy(x).pin_memory(), out=out_blk)
    assert relerr(out_blk.numpy(), (S @ x)[blk.r0 : blk.r1]) < 1e-12
    # caller-owned symmetric out buffer: the kernel stores into every rank's copy of it, no staging copy
    yr = dist.replicated_empty(n, torch.float64)
    xt = torch.from_numpy(x).cuda()
    for s_ in (1.0, -2.0, 0.5):
        got = A.dot(s_ * xt, out=yr)
        assert got is yr and relerr(yr.cpu().numpy(), s_ * (S @ x)) < 1e-12
    # several calls in a row: the replicated result buffers alternate (no opening barrier)
    for s_ in (1.0, 2.0, 3.0, 4.0):
        assert relerr(A @ (s_ * x), s_ * (S @ x)) < 1e-12
    # --- in-kernel all-reduce through the peer-mapped boards: deterministic, same on every rank
    board = dist.scalar_board()
    if board is not None:
        for it in range(6):
            t = torch.tensor([float(rank + 1) * 0.1 + it], dtype=torch.float64, device="cuda")
            board.allreduce(t, it % 3)
            want = sum((g + 1) * 0.1 + it for g in range(G))     # summed in rank order, like the kernel
            acc = 0.0
            for g in range(G):
                acc += float((g + 1) * 0.1 + it)
            assert t.item() == acc, (t.item(), acc, want)
        cur, prev = torch.tensor([5.0], device="cuda", dtype=torch.float64), torch.zeros(1, device="cuda", dtype=torch.float64)
        t = torch.tensor([1.0], dtype=torch.float64, device="cuda")
        board.allreduce(t, 2, cur_out=cur, prev_out=prev)
        assert prev.item() == 5.0 and cur.item() == float(G) and t.item() == float(G)
        board.check()
    # --- GMRES on row-sharded vectors (local CGS kernels + small all-reduces) vs the oracle / scipy residual
    ng = 3000
    Wg = sp.random(ng, ng, density=0.002, format="csr", random_state=21) + sp.eye(ng, format="csr") * 4.0
    Wg = Wg.tocsr()
    bg = rng.standard_normal(ng)
    xg, info = linalg.gmres(sparse.csr_array(Wg), bg, rtol=1e-10, restart=20, maxiter=400)
    assert info == 0 and relerr(Wg @ xg, bg) < 1e-9
    xo_g, _ = oracle.gmres(lambda v: Wg @ v, bg, rtol=1e-10, restart=20, maxiter=400)
    assert relerr(xg, xo_g) < 1e-8
    os.environ["LEGATE_SPARSE_GMRES_REPLICATED"] = "1"
    xr, _ = linalg.gmres(sparse.csr_array(Wg), bg, rtol=1e-10, restart=20, maxiter=400)
    os.environ.pop("LEGATE_SPARSE_GMRES_REPLICATED")
    assert relerr(xg, xr) < 1e-9
    torch.cuda.synchronize()
    print(f"rank {rank}/{G} OK", flush=True)
    dist.shutdown()


if __name__ == "__main__":
    main()
