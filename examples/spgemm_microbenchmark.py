#!/usr/bin/env python
"""SpGEMM micro-benchmark (protocol of the reference's examples/spgemm_microbenchmark.py:
banded A @ A.copy(), or MatrixMarket inputs, 5 warm-ups, ms/iteration printed)."""
import argparse

from _common import CudaTimer, banded_csr, parse_size, pick_package

if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("-n", "--nrows", default="1k")
    ap.add_argument("--nnz-per-row", type=int, default=11)
    ap.add_argument("-f", "--file", default="")
    ap.add_argument("-i", "--iters", type=int, default=10)
    ap.add_argument("--package", default="b200", choices=["b200", "scipy"])
    args = ap.parse_args()
    sparse, _, gpu = pick_package(args.package)
    if args.file:
        if gpu:
            A = sparse.mmread(args.file)
        else:
            import scipy.io

            A = sparse.csr_array(scipy.io.mmread(args.file))
    else:
        A = banded_csr(sparse, parse_size(args.nrows), args.nnz_per_row)
    B = A.copy()
    for _ in range(5):
        C = A @ B
    t = CudaTimer(gpu)
    t.start()
    for _ in range(args.iters):
        C = A @ B
    ms = t.stop() / args.iters
    print(f"SPGEMM {A.shape}x{B.shape} , nnz ({A.nnz})x({B.nnz})->({C.nnz}) : ms / iteration: {ms:.3f}")
