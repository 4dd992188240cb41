"""Phase cycle counters of the products consumer (debug library built with -DB2S_PIPE_TIMING, see
DESIGN §3.1): group 0 / thread 0 of CTA 0, summed over its tiles.  B2S_LIBRARY must point at the
instrumented .so."""
import ctypes, os, sys
ROOT = os.path.normpath(os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
sys.path[:0] = [ROOT, os.path.join(ROOT, "legate-sparse_b200"), os.path.join(ROOT, "tools")]
import numpy as np, torch
import legate_sparse as sparse
from legate_sparse import _native as N
import bench
lib = N.load()
lib.b2s_debug_pipe_phases.argtypes = [ctypes.c_void_p]
names = ["loop top + y prefetch", "wait full (TMA data)", "meta / row-list build", "products: LDS, gathers, STS", "group barrier",
         "pass 1: short rows", "pass 2: long rows", "(unused)"]
def report(tag, A, x, reps=5):
    y = A @ x
    buf = (ctypes.c_ulonglong * 16)()
    lib.b2s_debug_pipe_phases(buf)
    for _ in range(reps):
        y = A @ x
    lib.b2s_debug_pipe_phases(buf)
    v = np.array(list(buf)[:7], dtype=np.float64) / reps
    print(f"== {tag}: cycles per SpMV in group 0 of CTA 0 (total {v.sum():.0f})")
    for n_, c in zip(names, v):
        print(f"   {n_:32s} {c:12.0f}  {100 * c / v.sum():5.1f} %")
dev = torch.device("cuda")
vals, cols, ptr, x, nnz = bench.powerlaw_matrix(8_000_000, dev)
report("power-law 8M (long-row instance)", sparse.csr_array((vals, cols, ptr), shape=(8_000_000, 8_000_000)), x)
del vals, cols, ptr
n, k = 10_000_000, 50
A = sparse.random(n, n, density=k / n, rng=1234)
report("random C2 (column-blocked, both launches)", A, torch.rand(n, dtype=torch.float64, device=dev))
