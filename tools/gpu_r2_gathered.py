"""gathered SpMV (public A @ x with y replicated) at N ranks: unicast P2P stores vs NVSwitch multicast vs NCCL"""
import os, sys, time
ROOT = os.path.normpath(os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
sys.path[:0] = [ROOT, os.path.join(ROOT, "legate-sparse_b200")]
import numpy as np, torch
import legate_sparse as sparse
from legate_sparse import dist
import torch.distributed as td
dist.init()
G, rank = dist.world_size(), dist.rank()
dev = torch.device("cuda", torch.cuda.current_device())
n, k = 10_000_000, 50
A = sparse.random(n, n, density=k / n, rng=1234)
x = torch.rand(n, dtype=torch.float64, device=dev)
y_full = torch.empty(n, dtype=torch.float64, device=dev)
blk = A._block()
y_loc = torch.empty(blk.nrows, dtype=torch.float64, device=dev)

def timeit(fn, steps=20, warm=5):
    for _ in range(warm): fn()
    torch.cuda.synchronize(); td.barrier(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps): fn()
    e1.record(); torch.cuda.synchronize(); td.barrier()
    t = torch.tensor([e0.elapsed_time(e1) / steps], device=dev, dtype=torch.float64)
    td.all_reduce(t, op=td.ReduceOp.MAX)
    return float(t.item())

res = {"sharded": timeit(lambda: A.dot_local(x, out=y_loc))}
res["gathered_out"] = timeit(lambda: A.dot(x, out=y_full))
ref = y_full.clone()
res["gathered_clone"] = timeit(lambda: A.dot(x))
y_sym = dist.replicated_empty(n, torch.float64)
res["gathered_symm_out"] = timeit(lambda: A.dot(x, out=y_sym))
assert torch.equal(y_sym, ref)
# barrier cost alone
sv = dist.symm_vector(n, torch.float64, "spmv_y0")
if sv is not None:
    res["symm_barrier"] = timeit(lambda: sv.barrier(), steps=50)
    bd = dist.scalar_board()
    tok = torch.zeros(1, dtype=torch.float64, device=dev)
    res["board_exchange"] = timeit(lambda: bd.allreduce(tok, 0), steps=50)
res["nccl_allgather_80MB"] = timeit(lambda: dist.allgather_into(y_full, A.row_bounds()))
res["copy_80MB"] = timeit(lambda: y_full.copy_(ref))
ok = bool(torch.equal(ref[blk.r0:blk.r1], y_loc))
if rank == 0:
    print(os.environ.get("LEGATE_SPARSE_MULTICAST", "0"), G, {k: round(v, 4) for k, v in res.items()}, "own block exact:", ok, flush=True)
dist.shutdown()
