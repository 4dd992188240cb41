"""e2e host-vector SpMV (N=1): row-chunk count of the 2-D pipeline"""
import os, sys, time
ROOT = os.path.normpath(os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
sys.path[:0] = [ROOT, os.path.join(ROOT, "legate-sparse_b200")]
import numpy as np, torch
import legate_sparse as sparse
n, k = 10_000_000, 50
A = sparse.random(n, n, density=k / n, rng=1234)
x_host = torch.rand(n, dtype=torch.float64).pin_memory()
y_host = torch.empty(n, dtype=torch.float64).pin_memory()
for chunks in (2, 4, 6, 8, 12, 16):
    os.environ["LEGATE_SPARSE_HOSTPIPE_CHUNKS"] = str(chunks)
    A._block().hostpipe = None
    for _ in range(3):
        A.dot(x_host, out=y_host)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(10):
        A.dot(x_host, out=y_host)
    torch.cuda.synchronize()
    ms = (time.perf_counter() - t0) / 10 * 1e3
    print(f"chunks={chunks:2d}  {ms:.3f} ms  {2.0 * n * k / ms / 1e6:.1f} GFLOP/s", flush=True)
# copy-only floor: H2D x and D2H y concurrently
xd = torch.empty(n, dtype=torch.float64, device="cuda"); yd = torch.empty(n, dtype=torch.float64, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
torch.cuda.synchronize(); t0 = time.perf_counter()
for _ in range(10):
    with torch.cuda.stream(s1): xd.copy_(x_host, non_blocking=True)
    with torch.cuda.stream(s2): y_host.copy_(yd, non_blocking=True)
torch.cuda.synchronize()
print(f"copy-only (H2D 80 MB || D2H 80 MB): {(time.perf_counter() - t0) / 10 * 1e3:.3f} ms")
torch.cuda.synchronize(); t0 = time.perf_counter()
for _ in range(10):
    xd.copy_(x_host, non_blocking=True)
torch.cuda.synchronize()
print(f"H2D 80 MB alone: {(time.perf_counter() - t0) / 10 * 1e3:.3f} ms")
