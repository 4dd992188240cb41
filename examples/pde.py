#!/usr/bin/env python
"""CG on the Dirichlet Poisson problem (what the reference's examples/pde.py measures):
solve A u = b on an N x N grid, report iterations, time and ms/iteration; --throughput runs a
fixed number of iterations without a convergence exit."""
import argparse
import time

from _common import pick_package, poisson2d

import numpy as np

if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("-n", "--grid", type=int, default=256)
    ap.add_argument("-m", "--max-iters", type=int, default=None)
    ap.add_argument("--rtol", type=float, default=1e-10)
    ap.add_argument("--throughput", action="store_true")
    ap.add_argument("--package", default="b200", choices=["b200", "scipy"])
    args = ap.parse_args()
    sparse, linalg, gpu = pick_package(args.package)
    A = poisson2d(sparse, args.grid)
    b = np.random.default_rng(0).random(args.grid**2)
    if gpu:
        import torch

        b = torch.from_numpy(b).cuda()
    t0 = time.perf_counter()
    if args.throughput:
        iters = args.max_iters or 300
        x, it = linalg.cg(A, b, rtol=0.0, atol=0.0, maxiter=iters) if gpu else linalg.cg(A, b, rtol=0.0, maxiter=iters)
        it = iters
    elif gpu:
        x, it = linalg.cg(A, b, rtol=args.rtol, maxiter=args.max_iters)
    else:
        count = [0]
        x, _ = linalg.cg(A, b, rtol=args.rtol, maxiter=args.max_iters, callback=lambda _: count.__setitem__(0, count[0] + 1))
        it = count[0]
    if gpu:
        torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    res = b - A @ x
    rel = float(res.norm() / b.norm()) if gpu else float(np.linalg.norm(res) / np.linalg.norm(b))
    print(f"grid {args.grid}x{args.grid}: {it} CG iterations in {dt * 1e3:.1f} ms ({dt / max(it, 1) * 1e3:.4f} ms/iter), "
          f"relative residual {rel:.3e}")
