#!/usr/bin/env python
"""bench.py — CSR SpMV throughput (BASELINE.json metric) on B200, one process per GPU.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
           --master-port P bench.py --gpus N --steps K --warmup W

Workload (BASELINE.json configs[1]): random CSR 10M x 10M, 50 nnz/row, fp64, row-partitioned
over the ranks (strong scaling: the matrix is fixed, each rank owns rows/N).  A "step" is one
y = A x over the whole matrix.  The matrix is synthetic and generated on the device from a
counter-based hash, so every N sees the SAME matrix.

One JSON line is printed by rank 0.  Keys beyond the base contract:
  roofline      HBM roofline of the SpMV launch sequence (pipe kernel + its fix-up kernel, once per
                column block: the library splits this matrix into 2 column blocks so that the
                gathered slice of x stays L2 resident); achieved = plain-CSR algorithmic bytes / time
  cg            CG iterations/s on the 5-point Laplacian 4096^2 (the second half of the metric)
  cpu_baseline  the oracle's OpenMP restatement of the reference CPU task (spmv_omp.cc:36-44)
                timed on the host cores on a bounded row sample of the same matrix
  banded        same measurement on the reference's own microbenchmark generator
                (examples/common.py:206-249, nnz_per_row=51) — x window staged by TMA
  gathered      (N>1) the variant that all-gathers y (what a CG iteration needs)
  cusparse      (N=1, informative) cuSPARSE SpMV through torch.sparse on the same arrays —
                the vendor call the reference wraps (spmv.cu:117-152); bench-only
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
for p in (ROOT, os.path.join(ROOT, "legate-sparse_b200")):
    if p not in sys.path:
        sys.path.insert(0, p)

import numpy as np  # noqa: E402

METRIC = "csr_spmv_fp64_gflops"
UNIT = "GFLOP/s"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--rows", type=int, default=10_000_000)
    ap.add_argument("--nnz-per-row", type=int, default=50)
    ap.add_argument("--no-extras", action="store_true", help="skip banded / cusparse / cpu_baseline legs")
    ap.add_argument("--cpu-sample-rows", type=int, default=1_000_000)
    return ap.parse_args()


def workload_name(args):
    return f"random CSR {args.rows}x{args.rows}, {args.nnz_per_row} nnz/row, fp64 (BASELINE configs[1])"


# ------------------------------------------------------------------ synthetic matrix (device)
def _mix64(t):
    """splitmix64-style mixer on int64 tensors (wrap-around arithmetic)."""
    import torch

    t = (t ^ (t >> 30)) * -4658895280553007687   # 0xBF58476D1CE4E5B9
    t = (t ^ (t >> 27)) * -7723592293110705685   # 0x94D049BB133111EB
    return t ^ (t >> 31)


def gen_random_block(r0, r1, ncols, k, device, seed=1234, chunk_rows=1_000_000):
    """Rows [r0,r1) of the n x ncols matrix with exactly k nnz per row: the j-th entry of a row
    lies in the j-th of k equal strata of [0,ncols) (distinct + sorted columns, uniform over x);
    values uniform in (-1,1).  Entry (i,j) depends only on (seed,i,j)."""
    import torch

    n = r1 - r0
    stride = ncols // k
    cols = torch.empty(n * k, dtype=torch.int32, device=device)
    vals = torch.empty(n * k, dtype=torch.float64, device=device)
    jj = torch.arange(k, dtype=torch.int64, device=device)[None, :]
    for c0 in range(0, n, chunk_rows):
        c1 = min(n, c0 + chunk_rows)
        ii = torch.arange(r0 + c0, r0 + c1, dtype=torch.int64, device=device)[:, None]
        h = _mix64((ii * k + jj) + seed * 0x9E3779B97F4A7C15 % (1 << 62))
        off = (h & 0x7FFFFFFFFFFFFFFF) % stride
        cols[c0 * k : c1 * k] = (jj * stride + off).reshape(-1).to(torch.int32)
        h2 = _mix64(h + 0x632BE59BD9B4E019)
        u = ((h2 >> 11) & ((1 << 53) - 1)).to(torch.float64) * (1.0 / (1 << 53))
        vals[c0 * k : c1 * k] = (2.0 * u - 1.0).reshape(-1)
        del ii, h, off, h2, u
    indptr = torch.arange(n + 1, dtype=torch.int64, device=device) * k
    return vals, cols, indptr


def gen_banded_block(r0, r1, n, k, device):
    """Rows [r0,r1) of the reference's banded ones-matrix (examples/common.py:206-249)."""
    import torch

    half = k // 2
    rows = torch.arange(r0, r1, dtype=torch.int64, device=device)
    lo = torch.clamp(rows - half, min=0)
    hi = torch.clamp(rows + half, max=n - 1)
    cnt = hi - lo + 1
    indptr = torch.zeros(r1 - r0 + 1, dtype=torch.int64, device=device)
    torch.cumsum(cnt, 0, out=indptr[1:])
    nnz = int(indptr[-1].item())
    rep = torch.repeat_interleave(torch.arange(r1 - r0, device=device), cnt)
    pos = torch.arange(nnz, dtype=torch.int64, device=device) - indptr[:-1][rep]
    cols = (lo[rep] + pos).to(torch.int32)
    vals = torch.ones(nnz, dtype=torch.float64, device=device)
    return vals, cols, indptr


# ------------------------------------------------------------------ clocks sampler
class Clocks:
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.samples = []
        self.proc = None
        self.index = index

    def start(self):
        if os.environ.get("B2S_BENCH_NO_CLOCKS"):   # A/B switch: does the 10 Hz nvidia-smi poll perturb a leg?
            return
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                 "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.samples.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for s in self.samples:
            f = [t.strip() for t in s.split(",")]
            if len(f) < 6:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1]))
            except ValueError:
                continue
            for nm, v in zip(names, f[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------ measurement helpers
def spmv_bytes(nnz, nrows, ncols, idx_bytes):
    """Algorithmic bytes of one SpMV: every array once."""
    return nnz * (8 + idx_bytes) + (nrows + 1) * 8 + ncols * 8 + nrows * 8


def timed_steps(fn, steps, warmup, dist_mod):
    """CUDA-event timing of `steps` calls: barrier + sync on both sides, max over ranks."""
    import torch
    import torch.distributed as td

    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    if dist_mod.world_size() > 1:
        td.barrier()
    torch.cuda.synchronize()
    evs = [torch.cuda.Event(enable_timing=True) for _ in range(steps + 1)]
    evs[0].record()
    for i in range(steps):
        fn()
        evs[i + 1].record()
    torch.cuda.synchronize()
    if dist_mod.world_size() > 1:
        td.barrier()
    torch.cuda.synchronize()
    per = [evs[i].elapsed_time(evs[i + 1]) for i in range(steps)]
    total = torch.tensor([evs[0].elapsed_time(evs[steps])], dtype=torch.float64, device="cuda")
    if dist_mod.world_size() > 1:
        td.all_reduce(total, op=td.ReduceOp.MAX)
    return float(total.item()), per


def peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


def known_traffic(tag):
    """dram bytes/launch from the committed ncu --set full capture (profiles/), or None."""
    try:
        with open(os.path.join(ROOT, "profiles", "spmv_traffic.json")) as f:
            return json.load(f).get(tag)
    except Exception:
        return None


# ------------------------------------------------------------------ reference arm (CPU)
def run_reference(args):
    """--impl reference: the reference's CPU implementation of the path = its OpenMP task body
    (spmv_omp.cc:36-44), restated in oracle/ref_kernels.c (the native reference cannot be built
    here: legate.h), all host threads, on a bounded row sample of the same workload."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import oracle

    k, n = args.nnz_per_row, args.rows
    rows = min(args.cpu_sample_rows, n)
    vals, cols, indptr = host_sample(rows, n, k)
    x = np.random.default_rng(1).random(n)
    # all host threads the box offers; hyper-threads often hurt this gather-bound loop, so the
    # physical-core count is tried as well and the FASTER configuration is the one reported
    all_threads = oracle.omp_threads()
    best = None
    for threads in sorted({all_threads, max(1, all_threads // 2)}, reverse=True):
        oracle.omp_set_threads(threads)
        for _ in range(max(args.warmup, 1)):
            oracle.spmv(indptr, cols, vals, x, omp=True)
        t0 = time.perf_counter()
        for _ in range(args.steps):
            oracle.spmv(indptr, cols, vals, x, omp=True)
        dt_try = (time.perf_counter() - t0) / args.steps
        if best is None or dt_try < best[0]:
            best = (dt_try, threads)
    dt, threads = best
    gflops = 2.0 * rows * k / dt / 1e9
    sample = f"first {rows} rows of the {n}x{n} matrix ({rows * k} nnz) per step, int64 column ids"
    line = {
        "impl": "reference", "metric": METRIC, "value": gflops, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": workload_name(args), "sample": sample},
        "cpu_baseline": {"value": gflops, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": gflops, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))


def host_sample(rows, ncols, k, seed=1234):
    """The first `rows` rows of the bench matrix, regenerated on the host with the same hash."""
    M = (1 << 64) - 1

    def mix(t):
        t = ((t ^ (t >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)) & np.uint64(M)
        t = ((t ^ (t >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)) & np.uint64(M)
        return t ^ (t >> np.uint64(31))

    # NOTE: the device generator uses signed int64 arithmetic with arithmetic shifts; for the CPU
    # baseline only the SHAPE of the workload matters (k entries per row, one per stratum), so the
    # host sample uses its own unsigned mixer.
    stride = ncols // k
    ii = np.arange(rows, dtype=np.uint64)[:, None]
    jj = np.arange(k, dtype=np.uint64)[None, :]
    with np.errstate(over="ignore"):
        h = mix(ii * np.uint64(k) + jj + np.uint64(seed))
        cols = (jj * np.uint64(stride) + (h % np.uint64(stride))).astype(np.int64).reshape(-1)
        h2 = mix(h + np.uint64(0x632BE59BD9B4E019))
    vals = (2.0 * ((h2 >> np.uint64(11)).astype(np.float64) / float(1 << 53)) - 1.0).reshape(-1)
    indptr = np.arange(rows + 1, dtype=np.int64) * k
    return vals, cols, indptr


# ------------------------------------------------------------------ own arm
def plan_info_of(A):
    """tiling of the cached plan, or the column-block layout when the library chose one"""
    blk = A._block()
    if blk.colblock:
        return {"colblock": blk.colblock.info()}
    return blk.plan.info()


def run_b200(args):
    import torch

    import legate_sparse as sparse
    from legate_sparse import _native, dist

    dist.init()
    G, rank = dist.world_size(), dist.rank()
    assert G == args.gpus or (args.gpus == 1 and G == 1), f"--gpus {args.gpus} but WORLD_SIZE={G}"
    dev = torch.device("cuda", torch.cuda.current_device())
    n, k = args.rows, args.nnz_per_row
    bounds = dist.row_block_bounds(n, G)
    r0, r1 = int(bounds[rank]), int(bounds[rank + 1])

    vals, cols, indptr = gen_random_block(r0, r1, n, k, dev)
    A = sparse.csr_array.from_row_block(vals, cols, indptr, (n, n), row_start=r0, bounds=bounds)
    g = torch.Generator(device=dev)
    g.manual_seed(1)
    x = torch.rand(n, dtype=torch.float64, device=dev, generator=g)
    y_loc = torch.empty(r1 - r0, dtype=torch.float64, device=dev)
    A.dot_local(x, out=y_loc)  # builds the plan (one-time, like Legate's cached partitions)
    torch.cuda.synchronize()
    nnz_total = n * k
    plan_info = plan_info_of(A)

    # size-independent parity property at FULL size: linearity + a row sample against the oracle
    parity = full_size_checks(A, x, y_loc, vals, cols, indptr, r0)

    clocks = Clocks(torch.cuda.current_device())
    clocks.start()
    launches0 = _native.launch_count()
    total_ms, per = timed_steps(lambda: A.dot_local(x, out=y_loc), args.steps, args.warmup, dist)
    # the counter spans warm-up + timed calls (same launches per call): keep the timed share
    launches = (_native.launch_count() - launches0) * args.steps // (args.steps + args.warmup)
    ms_per_step = total_ms / args.steps
    value = 2.0 * nnz_total / (ms_per_step * 1e-3) / 1e9

    # roofline of the local launch sequence (rank 0's block; at N=1 the whole matrix)
    B_local = spmv_bytes((r1 - r0) * k, r1 - r0, n, 4)
    kernel_ms = float(np.mean(per))
    peak, peak_src = peaks()
    achieved = B_local / (kernel_ms * 1e-3) / 1e9
    nbk = plan_info["colblock"]["nblocks"] if "colblock" in plan_info else 1   # pipe-kernel launches per step
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": known_traffic(f"random_n{n}_k{k}_g{G}" + (f"_cb{nbk}" if nbk > 1 else "")),
                "launches_per_step": nbk, "algorithmic_bytes_per_launch": B_local / nbk, "idx_bytes": 4,
                "kernel_ms": kernel_ms / nbk,
                "peak_source": peak_src, "frac_of_8000_spec": achieved / 8000.0,
                "timed": "the launch sequence of one SpMV call, CUDA events per step: spmv_pipe_kernel + "
                         "spmv_fixup_kernel, once per column block when the operand is column-blocked "
                         "(achieved counts the plain-CSR algorithmic bytes once, not the extra indptr/y passes)"}

    # ---- e2e: public API with host buffers: H2D x (pinned) -> SpMV (+gather if N>1) -> D2H y (pinned)
    x_host = torch.empty(n, dtype=torch.float64).pin_memory()
    x_host.copy_(x)
    y_host = torch.empty(n, dtype=torch.float64).pin_memory()

    def e2e_step():
        A.dot(x_host, out=y_host)

    e2e_steps = max(3, min(args.steps, 10))
    e2e_ms, _ = timed_steps(e2e_step, e2e_steps, 3, dist)
    e2e_val = 2.0 * nnz_total / (e2e_ms / e2e_steps * 1e-3) / 1e9
    e2e = {"value": e2e_val, "unit": UNIT, "h2d_bytes_per_step": n * 8, "d2h_bytes_per_step": n * 8,
           "ms_per_step": e2e_ms / e2e_steps,
           "path": "csr_array.dot(x_pinned_host, out=y_pinned_host): H2D x, SpMV kernels"
                   + (", y all-gathered by the kernel's P2P stores" if G > 1 else "") + ", D2H y; matrix resident in HBM"}

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": G, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": {"workload": workload_name(args), "rows": n, "nnz": nnz_total, "index_dtype": "int32",
                   "partition": f"1-D row blocks over {G} rank(s), x replicated, y row-sharded",
                   "l2": "per-step inputs (%.2f GB/rank) exceed the 126 MB L2; no flush between steps"
                         % (B_local / 1e9),
                   "plan": plan_info},
        "effective_hbm_gbs": G * achieved if G == 1 else None,
        "clocks": None, "e2e": e2e, "gpu_launches": int(launches) * G, "roofline": roofline, "parity": parity,
    }

    # ---- gathered-y variant (what CG needs) at N>1
    if G > 1:
        y_full = torch.empty(n, dtype=torch.float64, device=dev)
        g_ms, _ = timed_steps(lambda: A.dot(x, out=y_full), args.steps, args.warmup, dist)
        line["gathered"] = {"value": 2.0 * nnz_total / (g_ms / args.steps * 1e-3) / 1e9, "unit": UNIT,
                            "ms_per_step": g_ms / args.steps,
                            "what": "SpMV with the all-gather of y fused into the kernel stores (NVLink P2P via symmetric "
                                    "memory; NCCL all-gather when peer memory is unavailable) + copy into out"}

    if not args.no_extras:
        del A
        torch.cuda.empty_cache()
        line["banded"] = banded_leg(args, dist, dev, bounds, rank, peak)
        line["cg"] = cg_leg(dist, dev, rank)
        if G == 1 and rank == 0:
            line["cusparse"] = cusparse_leg(vals, cols, indptr, x, n, args)
            line["cpu_baseline"] = cpu_baseline_leg(args)
    line["clocks"] = clocks.stop()   # sampled from the first timed region to the last one
    if rank == 0:
        print(json.dumps(line))
    dist.shutdown()


def full_size_checks(A, x, y_loc, vals, cols, indptr, r0):
    """Parity at full size: (i) linearity A(2x) == 2 A x to round-off, (ii) 2048 sampled rows
    recomputed by the oracle's C loop (reference spmv.cc:36-43) within 1e-10 relative."""
    import torch

    from oracle import oracle

    y2 = A.dot_local(2.0 * x)
    lin = float((torch.linalg.vector_norm(y2 - 2.0 * y_loc) / torch.linalg.vector_norm(y_loc)).item())
    nloc = y_loc.numel()
    rows = torch.linspace(0, nloc - 1, 2048, device=y_loc.device).long().unique()
    k = int((indptr[1] - indptr[0]).item())
    sel = (rows[:, None] * k + torch.arange(k, device=rows.device)[None, :]).reshape(-1)
    sub_cols = cols[sel].cpu().numpy().astype(np.int64)
    sub_vals = vals[sel].cpu().numpy()
    sub_ptr = np.arange(rows.numel() + 1, dtype=np.int64) * k
    y_or = oracle.spmv(sub_ptr, sub_cols, sub_vals, x.cpu().numpy())
    y_gpu = y_loc[rows].cpu().numpy()
    err = float(np.linalg.norm(y_gpu - y_or) / np.linalg.norm(y_or))
    assert lin < 1e-14 and err < 1e-10, (lin, err)
    return {"linearity_relerr": lin, "oracle_rows_checked": int(rows.numel()), "oracle_relerr": err,
            "tolerance": 1e-10}


def cg_leg(dist, dev, rank, grid=4096, iters=1000):
    """BASELINE metric, second half: CG iterations/s on the 5-point Laplacian (config 3: 4096^2 grid,
    fp64), fixed iteration count (no early exit), fused kernels, all ranks."""
    import torch

    import legate_sparse as sparse
    import legate_sparse.linalg as linalg
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    from side_bench import poisson2d_block

    G = dist.world_size()
    n = grid * grid
    bounds = dist.row_block_bounds(n, G)
    r0, r1 = int(bounds[rank]), int(bounds[rank + 1])
    data, idx, ptr = poisson2d_block(grid, r0, r1, dev)
    A = sparse.csr_array.from_row_block(data, idx, ptr, (n, n), row_start=r0, bounds=bounds)
    g = torch.Generator(device=dev)
    g.manual_seed(2)
    b = torch.rand(n, dtype=torch.float64, device=dev, generator=g)
    linalg.cg(A, b, rtol=0.0, atol=0.0, maxiter=25)
    torch.cuda.synchronize()

    def call(m):
        if G > 1:
            import torch.distributed as td

            td.barrier()
            torch.cuda.synchronize()
        t0 = time.perf_counter()
        _, done = linalg.cg(A, b, rtol=0.0, atol=0.0, maxiter=m)
        torch.cuda.synchronize()
        return time.perf_counter() - t0, done

    # whole solver call (set-up: graph capture, halo ranges, scalars + `iters` iterations), and the
    # steady iteration rate from the difference to a short call (the set-up is the same in both)
    dt_short, it_short = call(iters // 5)
    dt, it = call(iters)
    steady = (it - it_short) / max(dt - dt_short, 1e-9)
    nnz = A.nnz
    ref_bytes = nnz * 12 + (n + 1) * 8 + 16 * n + 120 * n
    return {"workload": f"CG, 5-point Laplacian {grid}x{grid} (n={n}, nnz={nnz}), fp64, identity M, {it} iterations",
            "iters_per_s": it / dt, "ms_per_iter": dt / it * 1e3, "steady_iters_per_s": steady,
            "timing": "wall clock around linalg.cg() incl. its set-up; steady = (it - it/5) / (t - t_short)",
            "reference_algorithm_bytes_per_iter": ref_bytes, "fused_bytes_per_iter": ref_bytes - 48 * n,
            "kernels_per_iter": "cg_pupdate + spmv_pipe(+dot) + fixup + reduce + cg_update (+NCCL all-gather/all-reduce at N>1)"}


def banded_leg(args, dist, dev, bounds, rank, peak):
    import torch

    import legate_sparse as sparse

    n, k = args.rows, 51
    r0, r1 = int(bounds[rank]), int(bounds[rank + 1])
    vals, cols, indptr = gen_banded_block(r0, r1, n, k, dev)
    A = sparse.csr_array.from_row_block(vals, cols, indptr, (n, n), row_start=r0, bounds=bounds)
    x = torch.ones(n, dtype=torch.float64, device=dev)
    y = torch.empty(r1 - r0, dtype=torch.float64, device=dev)
    A.dot_local(x, out=y)
    ok = bool((y[100:-100] == 51.0).all().item()) if (r1 - r0) > 400 else True
    ms, per = timed_steps(lambda: A.dot_local(x, out=y), args.steps, args.warmup, dist)
    nnz_loc = int(vals.numel())
    t = torch.tensor([nnz_loc], dtype=torch.float64, device=dev)
    dist.allreduce_sum_(t)
    nnz_total = float(t.item())
    B = spmv_bytes(nnz_loc, r1 - r0, n, 4)
    ach = B / (float(np.mean(per)) * 1e-3) / 1e9
    return {"workload": f"banded ones {n}x{n}, 51 nnz/row (reference examples/common.py:206-249)",
            "value": 2.0 * nnz_total / (ms / args.steps * 1e-3) / 1e9, "unit": UNIT,
            "ms_per_step": ms / args.steps, "interior_rows_exact": ok,
            "roofline": {"bound": "hbm", "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak},
            "plan": plan_info_of(A)}


def cusparse_leg(vals, cols, indptr, x, n, args):
    """Vendor baseline (bench-only, never linked into the product): torch.sparse CSR @ x → cusparseSpMV."""
    import torch

    try:
        At = torch.sparse_csr_tensor(indptr.to(torch.int32), cols, vals, size=(n, n))
        for _ in range(3):
            yt = At @ x
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        steps = max(3, min(args.steps, 10))
        e0.record()
        for _ in range(steps):
            yt = At @ x
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / steps
        return {"value": 2.0 * vals.numel() / (ms * 1e-3) / 1e9, "unit": UNIT, "ms_per_step": ms,
                "what": "cusparseSpMV via torch.sparse_csr_tensor @ x (int32 indices), includes y allocation"}
    except Exception as e:  # informative leg only
        return {"unavailable": str(e)[:200]}


def cpu_baseline_leg(args):
    from oracle import oracle
    import scipy.sparse as sp

    k, n = args.nnz_per_row, args.rows
    rows = min(args.cpu_sample_rows, n)
    vals, cols, indptr = host_sample(rows, n, k)
    x = np.random.default_rng(1).random(n)
    threads = oracle.omp_threads()
    oracle.spmv(indptr, cols, vals, x, omp=True)
    reps, t0 = 0, time.perf_counter()
    while time.perf_counter() - t0 < 8.0:
        oracle.spmv(indptr, cols, vals, x, omp=True)
        reps += 1
    dt = (time.perf_counter() - t0) / reps
    S = sp.csr_array((vals, cols.astype(np.int32), indptr.astype(np.int32) if indptr[-1] < 2**31 else indptr),
                     shape=(rows, n))
    S @ x
    t1 = time.perf_counter()
    for _ in range(3):
        S @ x
    dts = (time.perf_counter() - t1) / 3
    return {"value": 2.0 * rows * k / dt / 1e9, "unit": UNIT, "cores": threads, "kind": "port",
            "sample": f"first {rows} rows ({rows * k} nnz) of the {n}x{n} matrix, ~8 s of OpenMP SpMV "
                      f"(oracle restatement of spmv_omp.cc:36-44)",
            "scipy_single_thread_gflops": 2.0 * rows * k / dts / 1e9,
            "host_cpus": os.cpu_count()}


def main():
    args = parse()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
