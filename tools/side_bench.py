#!/usr/bin/env python
"""Secondary measurements (not the bench.py headline): CG iterations/s on the 5-point Laplacian
(BASELINE config 3) and SpGEMM A@A on banded / R-MAT matrices (config 4), single GPU or torchrun.

    python tools/side_bench.py cg [--grid 4096] [--iters 200] [--no-solve]
    python tools/side_bench.py gmres [--grid 4096] [--iters 100]
    python tools/side_bench.py spgemm [--scale 20] [--banded-n 4000000]
Prints one JSON object per measurement.
"""
import argparse
import json
import os
import sys
import time

ROOT = os.path.normpath(os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
for p in (ROOT, os.path.join(ROOT, "legate-sparse_b200")):
    if p not in sys.path:
        sys.path.insert(0, p)

import numpy as np
import torch

import legate_sparse as sparse
import legate_sparse.linalg as linalg
from legate_sparse import dist


def poisson2d_block(N, r0, r1, device):
    """Rows [r0,r1) of the 5-point Poisson matrix on an N x N grid — the same entries as
    reference examples/common.py:313-327 (diags → tocsr drops the explicit zeros at the grid-row
    boundaries), generated directly on the device."""
    n = N * N
    i = torch.arange(r0, r1, dtype=torch.int64, device=device)
    offs = torch.tensor([-N, -1, 0, 1, N], dtype=torch.int64, device=device)
    vals = torch.tensor([-1.0, -1.0, 4.0, -1.0, -1.0], dtype=torch.float64, device=device)
    cols = i[:, None] + offs[None, :]
    ok = (cols >= 0) & (cols < n)
    ok[:, 1] &= (i % N) != 0          # (i, i-1) is zero when i starts a grid row
    ok[:, 3] &= ((i + 1) % N) != 0    # (i, i+1) is zero when i ends a grid row
    cnt = ok.sum(dim=1)
    indptr = torch.zeros(r1 - r0 + 1, dtype=torch.int64, device=device)
    torch.cumsum(cnt, 0, out=indptr[1:])
    indices = cols[ok].to(torch.int32)
    data = vals[None, :].expand(r1 - r0, 5)[ok].contiguous()
    return data, indices, indptr


def rmat_device(scale, edge_factor=16, a=0.57, b=0.19, c=0.19, seed=42, device="cuda"):
    g = torch.Generator(device=device)
    g.manual_seed(seed)
    n = 1 << scale
    ne = edge_factor * n
    rows = torch.zeros(ne, dtype=torch.int64, device=device)
    cols = torch.zeros(ne, dtype=torch.int64, device=device)
    for _ in range(scale):
        r = torch.rand(ne, device=device, generator=g)
        rbit = (r >= a + b).to(torch.int64)
        cbit = (((r >= a) & (r < a + b)) | (r >= a + b + c)).to(torch.int64)
        rows = (rows << 1) | rbit
        cols = (cols << 1) | cbit
    key = torch.unique(rows * n + cols)          # sorted, duplicates removed (values = 1.0)
    rows, cols = key // n, key % n
    indptr = torch.zeros(n + 1, dtype=torch.int64, device=device)
    torch.cumsum(torch.bincount(rows, minlength=n), 0, out=indptr[1:])
    data = torch.ones(key.numel(), dtype=torch.float64, device=device)
    return data, cols.to(torch.int32), indptr, n


def run_cg(args):
    dist.init()
    G, rank = dist.world_size(), dist.rank()
    dev = torch.device("cuda", torch.cuda.current_device())
    N = args.grid
    n = N * N
    bounds = dist.row_block_bounds(n, G)
    r0, r1 = int(bounds[rank]), int(bounds[rank + 1])
    data, idx, ptr = poisson2d_block(N, r0, r1, dev)
    A = sparse.csr_array.from_row_block(data, idx, ptr, (n, n), row_start=r0, bounds=bounds)
    g = torch.Generator(device=dev)
    g.manual_seed(2)
    b = torch.rand(n, dtype=torch.float64, device=dev, generator=g)
    nnz = A.nnz
    out = {"what": f"CG on the 5-point Laplacian {N}x{N} grid (n={n}, nnz={nnz}), fp64, identity preconditioner",
           "n_gpus": G, "iters": args.iters}
    for mode in ("fused", "unfused"):
        os.environ["LEGATE_SPARSE_CG_UNFUSED"] = "1" if mode == "unfused" else "0"
        linalg.cg(A, b, rtol=0.0, atol=0.0, maxiter=25)  # warm-up (plan, allocator)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        x, it = linalg.cg(A, b, rtol=0.0, atol=0.0, maxiter=args.iters)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        res = float((torch.linalg.vector_norm(b - (A @ x)) / torch.linalg.vector_norm(b)).item())
        B = nnz * 12 + (n + 1) * 8 + 16 * n
        out[mode] = {"iters_per_s": it / dt, "ms_per_iter": dt / it * 1e3, "rel_residual_after": res,
                     "ref_algorithm_bytes_per_iter": B + 120 * n,
                     "ref_algorithm_gbs": (B + 120 * n) * it / dt / 1e9}
    os.environ["LEGATE_SPARSE_CG_UNFUSED"] = "0"
    # convergence run
    if not args.no_solve:
        t0 = time.perf_counter()
        x, it = linalg.cg(A, b, rtol=1e-10, maxiter=20000)
        torch.cuda.synchronize()
        out["solve_rtol_1e-10"] = {"iters": it, "seconds": time.perf_counter() - t0,
                                   "rel_residual": float((torch.linalg.vector_norm(b - (A @ x)) /
                                                          torch.linalg.vector_norm(b)).item())}
    if rank == 0:
        print(json.dumps(out))
    dist.shutdown()


def run_gmres(args):
    """restarted GMRES (restart 20) on the 5-point Laplacian: time per restart cycle and the HBM
    rate of its streams (SpMV B + CGS project (k+1)n + update (k+2)n + scale 2n values, k = 1..20)."""
    dist.init()
    G, rank = dist.world_size(), dist.rank()
    dev = torch.device("cuda", torch.cuda.current_device())
    N = args.grid
    n = N * N
    bounds = dist.row_block_bounds(n, G)
    r0, r1 = int(bounds[rank]), int(bounds[rank + 1])
    data, idx, ptr = poisson2d_block(N, r0, r1, dev)
    A = sparse.csr_array.from_row_block(data, idx, ptr, (n, n), row_start=r0, bounds=bounds)
    g = torch.Generator(device=dev)
    g.manual_seed(2)
    b = torch.rand(n, dtype=torch.float64, device=dev, generator=g)
    nnz = A.nnz
    restart, cycles = 20, args.iters // 20 if args.iters >= 20 else 5
    linalg.gmres(A, b, rtol=0.0, atol=0.0, restart=restart, maxiter=restart)   # warm-up
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    x, info = linalg.gmres(A, b, rtol=0.0, atol=0.0, restart=restart, maxiter=restart * cycles)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    B = nnz * 12 + (n + 1) * 8 + 16 * n
    per_cycle = restart * B + sum((2 * k + 5) for k in range(1, restart + 1)) * 8 * n + (restart + 2) * 8 * n + B + 24 * n
    res = float((torch.linalg.vector_norm(b - (A @ x)) / torch.linalg.vector_norm(b)).item())
    if rank == 0:
        print(json.dumps({"what": f"GMRES(20) on the 5-point Laplacian {N}x{N} (n={n}), fp64, {cycles} restart cycles",
                          "n_gpus": G, "ms_per_cycle": dt / cycles * 1e3, "iters_per_s": restart * cycles / dt,
                          "stream_bytes_per_cycle": per_cycle, "stream_gbs": per_cycle * cycles / dt / 1e9,
                          "rel_residual_after": res}))
    dist.shutdown()


def time_spgemm(A, reps=3):
    C = A @ A  # warm-up
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        C = A @ A
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / reps, C


def run_spgemm(args):
    dist.init()
    dev = torch.device("cuda", torch.cuda.current_device())
    res = []
    # banded A@A (what reference examples/spgemm_microbenchmark.py builds)
    for n, k in ((args.banded_n, 11),):
        half = k // 2
        rows = torch.arange(n, dtype=torch.int64, device=dev)
        lo = torch.clamp(rows - half, min=0)
        hi = torch.clamp(rows + half, max=n - 1)
        cnt = hi - lo + 1
        ptr = torch.zeros(n + 1, dtype=torch.int64, device=dev)
        torch.cumsum(cnt, 0, out=ptr[1:])
        rep = torch.repeat_interleave(torch.arange(n, device=dev), cnt)
        pos = torch.arange(int(ptr[-1]), dtype=torch.int64, device=dev) - ptr[:-1][rep]
        A = sparse.csr_array((torch.ones(int(ptr[-1]), dtype=torch.float64, device=dev),
                              (lo[rep] + pos).to(torch.int32), ptr), shape=(n, n))
        dt, C = time_spgemm(A)
        prod = C._last_products
        res.append({"what": f"banded {n}x{n}, {k}/row: A@A", "ms": dt * 1e3, "nnzA": A.nnz, "nnzC": C.nnz,
                    "products": prod, "gflops": 2.0 * prod / dt / 1e9,
                    "lower_bound_gbs": ((2 * A.nnz + C.nnz) * 12 + 3 * (n + 1) * 8) / dt / 1e9})
        try:   # vendor comparison (bench-only): cuSPARSE SpGEMM through torch.sparse.mm
            blk = A._block()
            At = torch.sparse_csr_tensor(blk.indptr.to(torch.int32), blk.indices, blk.data, size=(n, n))
            Ct = torch.sparse.mm(At, At)
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            for _ in range(3):
                Ct = torch.sparse.mm(At, At)
            torch.cuda.synchronize()
            res[-1]["cusparse_ms"] = (time.perf_counter() - t0) / 3 * 1e3
            del At, Ct
        except Exception as e:
            res[-1]["cusparse_error"] = str(e)[:160]
        del A, C
    for scale in args.scale:
        data, idx, ptr, n = rmat_device(scale, device=dev)
        A = sparse.csr_array((data, idx, ptr), shape=(n, n))
        torch.cuda.synchronize()
        try:
            dt, C = time_spgemm(A, reps=2)
            prod = C._last_products
            item = {"what": f"R-MAT scale {scale} (n={n}): A@A", "ms": dt * 1e3, "nnzA": int(data.numel()),
                    "nnzC": C.nnz, "products": prod, "gflops": 2.0 * prod / dt / 1e9,
                    "compression": prod / max(C.nnz, 1)}
            if scale <= args.verify_scale:
                import scipy.sparse as sp

                S = sp.csr_array((data.cpu().numpy(), idx.cpu().numpy(), ptr.cpu().numpy()), shape=(n, n))
                t0 = time.perf_counter()
                E = (S @ S).tocsr()
                item["scipy_ms"] = (time.perf_counter() - t0) * 1e3
                E.sort_indices()
                ok = np.array_equal(C.indptr, E.indptr) and np.array_equal(C.indices, E.indices)
                item["matches_scipy"] = bool(ok and np.allclose(C.data, E.data, rtol=1e-12))
            res.append(item)
        except RuntimeError as e:
            res.append({"what": f"R-MAT scale {scale}", "error": str(e)[:300]})
        del A
        torch.cuda.empty_cache()
    if dist.rank() == 0:
        for r in res:
            print(json.dumps(r))
    dist.shutdown()


def run_powerlaw(args):
    """BASELINE config 5: power-law row degrees (Zipf alpha=2 clipped to [1, 10000], one row at
    10000), n = 8M, uniform columns; SpMV through the public API vs cuSPARSE (torch.sparse)."""
    dist.init()
    dev = torch.device("cuda", torch.cuda.current_device())
    n = args.pl_rows
    g = torch.Generator(device=dev)
    g.manual_seed(7)
    u = torch.rand(n, device=dev, generator=g, dtype=torch.float64)
    # inverse-CDF sampling of a discrete power law P(d) ~ d^-2 on [1, 10000]
    deg = torch.clamp((1.0 / (1.0 - u * (1.0 - 1.0 / 10000.0))).floor().long(), 1, 10000)
    deg[n // 3] = 10000
    ptr = torch.zeros(n + 1, dtype=torch.int64, device=dev)
    torch.cumsum(deg, 0, out=ptr[1:])
    nnz = int(ptr[-1].item())
    cols = torch.randint(0, n, (nnz,), device=dev, generator=g, dtype=torch.int32)
    vals = torch.rand(nnz, device=dev, generator=g, dtype=torch.float64) - 0.5
    A = sparse.csr_array((vals, cols, ptr), shape=(n, n))
    x = torch.rand(n, device=dev, generator=g, dtype=torch.float64)
    y = A @ x
    # parity on a row sample (oracle C loop) incl. the longest row
    from oracle import oracle
    rows = torch.cat([torch.linspace(0, n - 1, 500, device=dev).long(), torch.tensor([n // 3], device=dev)])
    xs = x.cpu().numpy()
    worst = 0.0
    for r in rows.tolist():
        lo, hi = int(ptr[r]), int(ptr[r + 1])
        ref = oracle.spmv(np.array([0, hi - lo]), cols[lo:hi].cpu().numpy(), vals[lo:hi].cpu().numpy(), xs)[0]
        worst = max(worst, abs(float(y[r]) - ref) / max(abs(ref), 1e-300))
    out = {"what": f"power-law CSR n={n}, nnz={nnz}, max row 10000 (config 5)", "max_rel_err_vs_oracle_rows": worst}
    for name, fn in (("b200", lambda: A.dot(x, out=y)),):
        for _ in range(5):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20):
            fn()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 20
        out[name] = {"ms": ms, "gflops": 2.0 * nnz / ms / 1e6, "algorithmic_gbs": (nnz * 12 + n * 24) / ms / 1e6}
    try:
        At = torch.sparse_csr_tensor(ptr.to(torch.int32), cols, vals, size=(n, n))  # same index dtype
        for _ in range(3):
            At @ x
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            At @ x
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 10
        out["cusparse"] = {"ms": ms, "gflops": 2.0 * nnz / ms / 1e6}
    except Exception as e:
        out["cusparse"] = {"error": str(e)[:200]}
    print(json.dumps(out))
    # cuSPARSE SpGEMM comparison on R-MAT 16 and banded
    for scale in (16,):
        data, idx, p2, m = rmat_device(scale, device=dev)
        B = sparse.csr_array((data, idx, p2), shape=(m, m))
        dt, C = time_spgemm(B)
        item = {"what": f"R-MAT {scale} A@A", "b200_ms": dt * 1e3, "nnzC": C.nnz}
        try:
            Bt = torch.sparse_csr_tensor(p2, idx.long(), data, size=(m, m))
            Ct = torch.sparse.mm(Bt, Bt)
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            for _ in range(2):
                Ct = torch.sparse.mm(Bt, Bt)
            torch.cuda.synchronize()
            item["cusparse_ms"] = (time.perf_counter() - t0) / 2 * 1e3
            item["cusparse_nnzC"] = int(Ct._nnz())
        except Exception as e:
            item["cusparse_error"] = str(e)[:200]
        print(json.dumps(item))
    dist.shutdown()


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("which", choices=["cg", "gmres", "spgemm", "powerlaw"])
    ap.add_argument("--no-solve", action="store_true")
    ap.add_argument("--pl-rows", type=int, default=8_000_000)
    ap.add_argument("--grid", type=int, default=4096)
    ap.add_argument("--iters", type=int, default=200)
    ap.add_argument("--scale", type=int, nargs="*", default=[16, 18, 20])
    ap.add_argument("--verify-scale", type=int, default=16)
    ap.add_argument("--banded-n", type=int, default=4_000_000)
    a = ap.parse_args()
    {"cg": run_cg, "gmres": run_gmres, "spgemm": run_spgemm, "powerlaw": run_powerlaw}[a.which](a)
