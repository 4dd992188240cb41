"""Shared helpers for the examples (own implementation; the workloads follow the reference's
examples/common.py: banded ones-matrix :206-249, poisson2D :313-327)."""
import os
import sys
import time

ROOT = os.path.normpath(os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
for p in (ROOT, os.path.join(ROOT, "legate-sparse_b200")):
    if p not in sys.path:
        sys.path.insert(0, p)

import numpy as np  # noqa: E402


def parse_size(text: str) -> int:
    """'10m' → 10*1024*1024 like the reference CLI (examples/common.py:22-37)."""
    text = text.lower().strip()
    mult = {"k": 1024, "m": 1024**2, "g": 1024**3}.get(text[-1:], 1)
    return int(text[:-1] if text[-1:] in "kmg" else text) * mult


def banded_csr(sparse, n: int, nnz_per_row: int, dtype=np.float64):
    """n x n matrix of ones with `nnz_per_row` (odd) diagonals, built directly in CSR."""
    assert nnz_per_row % 2 == 1 and n > nnz_per_row
    half = nnz_per_row // 2
    rows = np.arange(n, dtype=np.int64)
    lo, hi = np.maximum(rows - half, 0), np.minimum(rows + half, n - 1)
    indptr = np.concatenate([[0], np.cumsum(hi - lo + 1)]).astype(np.int64)
    rep = np.repeat(rows, hi - lo + 1)
    cols = lo[rep] + (np.arange(int(indptr[-1]), dtype=np.int64) - indptr[rep])
    return sparse.csr_array((np.ones(int(indptr[-1]), dtype=dtype), cols.astype(np.int64), indptr), shape=(n, n))


def poisson2d(sparse, N: int):
    """5-point Laplacian on an N x N grid through sparse.diags (same diagonals as the reference)."""
    first = np.full(N - 1, -1.0)
    side = np.concatenate([first, np.tile(np.concatenate([[0.0], first]), (N * N - 1 - (N - 1)) // N)])
    far = -np.ones(N * (N - 1))
    return sparse.diags([far, side, 4.0 * np.ones(N * N), side, far], [-N, -1, 0, 1, N], dtype=np.float64).tocsr()


class CudaTimer:
    """ms between start() and stop() on the current CUDA stream (host timer for scipy)."""

    def __init__(self, use_cuda=True):
        self.use_cuda = use_cuda

    def start(self):
        if self.use_cuda:
            import torch

            self.e0, self.e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            self.e0.record()
        else:
            self.t0 = time.perf_counter()

    def stop(self) -> float:
        if self.use_cuda:
            import torch

            self.e1.record()
            torch.cuda.synchronize()
            return self.e0.elapsed_time(self.e1)
        return (time.perf_counter() - self.t0) * 1e3


def pick_package(name: str):
    """'b200' → this repo's legate_sparse (numpy/torch arrays), 'scipy' → scipy.sparse."""
    if name == "scipy":
        import scipy.sparse as sp
        import scipy.sparse.linalg as spla

        return sp, spla, False
    import legate_sparse as sp
    import legate_sparse.linalg as spla

    return sp, spla, True
