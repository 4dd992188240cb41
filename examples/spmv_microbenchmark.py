#!/usr/bin/env python
"""SpMV micro-benchmark with the protocol of the reference's examples/spmv_microbenchmark.py
(5 warm-ups, N timed iterations of y = A.dot(x, out=y) on a banded ones-matrix, ms/iter printed).

    python examples/spmv_microbenchmark.py --nmin 1m --nmax 8m --nnz-per-row 11 [--package scipy]
    torchrun --nproc-per-node 8 examples/spmv_microbenchmark.py --nmin 10m --nmax 10m
"""
import argparse

from _common import CudaTimer, banded_csr, parse_size, pick_package

import numpy as np

if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--nmin", default="1k")
    ap.add_argument("--nmax", default="1m")
    ap.add_argument("--nnz-per-row", type=int, default=11)
    ap.add_argument("-i", "--iters", type=int, default=100)
    ap.add_argument("--package", default="b200", choices=["b200", "scipy"])
    args = ap.parse_args()
    sparse, _, gpu = pick_package(args.package)
    if gpu:
        import torch
        from legate_sparse import dist

        dist.init()
    n = parse_size(args.nmin)
    while n <= parse_size(args.nmax):
        A = banded_csr(sparse, n, args.nnz_per_row)
        if gpu:
            x = torch.ones(n, dtype=torch.float64, device="cuda")
            y = torch.zeros(n, dtype=torch.float64, device="cuda")
            step = lambda: A.dot(x, out=y)  # noqa: E731
        else:
            x, y = np.ones(n), np.zeros(n)
            step = lambda: A @ x  # noqa: E731
        for _ in range(5):
            step()
        t = CudaTimer(gpu)
        t.start()
        for _ in range(args.iters):
            step()
        ms = t.stop() / args.iters
        if not gpu or dist.rank() == 0:
            print(f"SPMV rows: {n}, nnz: {A.nnz} , ms / iter: {ms:.4f}  ({2.0 * A.nnz / ms / 1e6:.1f} GFLOP/s)")
        n *= 2
