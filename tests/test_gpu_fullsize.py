"""Parity at BASELINE.json's FULL sizes through size-independent properties (the oracle cannot
run 5e8 non-zeros in seconds): linearity, an independent checksum of all products, oracle rows,
analytic row sums of the Laplacian, bit-reproducibility."""
import os
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.normpath(os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))


def test_config2_random_10m_50_per_row():
    import torch

    import bench
    import legate_sparse as sparse
    from oracle import oracle

    dev = torch.device("cuda")
    n, k = 10_000_000, 50
    vals, cols, indptr = bench.gen_random_block(0, n, n, k, dev)
    A = sparse.csr_array.from_row_block(vals, cols, indptr, (n, n))
    g = torch.Generator(device=dev)
    g.manual_seed(1)
    x = torch.rand(n, dtype=torch.float64, device=dev, generator=g)
    y = A.dot_local(x)
    # (1) bit-reproducible (no atomics anywhere on the path)
    assert torch.equal(y, A.dot_local(x))
    # (2) linearity: A(2x) == 2 A x exactly (scaling by 2 is exact in binary fp)
    assert torch.equal(A.dot_local(2.0 * x), 2.0 * y)
    # (3) checksum over ALL 5e8 products by an independent path (torch gather + sum, chunked)
    total = torch.zeros((), dtype=torch.float64, device=dev)
    absum = torch.zeros((), dtype=torch.float64, device=dev)
    step = 50_000_000
    for s in range(0, n * k, step):
        p = vals[s : s + step] * x[cols[s : s + step].long()]
        total += p.sum()
        absum += p.abs().sum()
    assert abs(float(y.sum() - total)) <= 1e-10 * float(absum)
    # (4) 4096 rows against the oracle's C loop (reference spmv.cc:36-43), 1e-10 relative
    rows = torch.linspace(0, n - 1, 4096, device=dev).long().unique()
    sel = (rows[:, None] * k + torch.arange(k, device=dev)[None, :]).reshape(-1)
    y_or = oracle.spmv(np.arange(rows.numel() + 1, dtype=np.int64) * k, cols[sel].cpu().numpy().astype(np.int64),
                       vals[sel].cpu().numpy(), x.cpu().numpy())
    y_g = y[rows].cpu().numpy()
    assert np.linalg.norm(y_g - y_or) / np.linalg.norm(y_or) < 1e-10
    # stratified generator: exactly k sorted distinct columns per row
    c = cols[: 1000 * k].view(1000, k).long()
    assert bool((c[:, 1:] > c[:, :-1]).all())


def test_config3_poisson_4096_row_sums_and_cg_step():
    import torch

    import legate_sparse as sparse
    import legate_sparse.linalg as linalg
    from side_bench import poisson2d_block

    dev = torch.device("cuda")
    N = 4096
    n = N * N
    data, idx, ptr = poisson2d_block(N, 0, n, dev)
    assert int(ptr[-1]) == 5 * n - 4 * N          # nnz of the 5-point stencil (83 869 696 at N=4096)
    A = sparse.csr_array.from_row_block(data, idx, ptr, (n, n))
    ones = torch.ones(n, dtype=torch.float64, device=dev)
    y = A.dot_local(ones)
    # analytic row sums: 4 - (number of neighbours inside the grid)
    i = torch.arange(n, device=dev)
    gi, gj = i // N, i % N
    neigh = (gi > 0).long() + (gi < N - 1).long() + (gj > 0).long() + (gj < N - 1).long()
    assert torch.equal(y, (4 - neigh).to(torch.float64))
    # symmetry property: <Ax, z> == <x, Az>
    g = torch.Generator(device=dev); g.manual_seed(3)
    xv = torch.rand(n, dtype=torch.float64, device=dev, generator=g)
    zv = torch.rand(n, dtype=torch.float64, device=dev, generator=g)
    lhs, rhs = torch.dot(A.dot_local(xv), zv), torch.dot(xv, A.dot_local(zv))
    assert abs(float(lhs - rhs)) <= 1e-12 * abs(float(lhs))
    # 50 CG iterations: fused kernels and the op-for-op reference sequence give the same iterate
    b = torch.rand(n, dtype=torch.float64, device=dev, generator=g)
    xf, itf = linalg.cg(A, b, rtol=0.0, atol=0.0, maxiter=50)
    os.environ["LEGATE_SPARSE_CG_UNFUSED"] = "1"
    try:
        xu, itu = linalg.cg(A, b, rtol=0.0, atol=0.0, maxiter=50)
    finally:
        os.environ["LEGATE_SPARSE_CG_UNFUSED"] = "0"
    assert itf == itu == 50
    assert float((xf - xu).norm() / xu.norm()) < 1e-10
