"""Seeded synthetic generators shared by tests and bench (SURVEY §8d inputs).
numpy/scipy only — usable on the CPU box and (small sizes) as oracle inputs."""
import numpy as np
import scipy.sparse as sp


def poisson2d_diagonals(N):
    """5-point Poisson diagonals exactly as reference examples/common.py:313-327."""
    diag_size = N * N - 1
    first = np.full((N - 1), -1.0)
    chunks = np.concatenate([np.zeros(1), first])
    diag_a = np.concatenate([first, np.tile(chunks, (diag_size - (N - 1)) // N)])
    diag_g = -1.0 * np.ones(N * (N - 1))
    diag_c = 4.0 * np.ones(N * N)
    return [diag_g, diag_a, diag_c, diag_a, diag_g], [-N, -1, 0, 1, N]


def poisson2d_scipy(N):
    d, o = poisson2d_diagonals(N)
    return sp.diags(d, o, dtype=np.float64).tocsr()


def banded_csr_arrays(N, nnz_per_row, dtype=np.float64, ones=True):
    """Banded matrix of the reference microbenchmark (examples/common.py:206-249), direct CSR."""
    assert N > nnz_per_row and nnz_per_row % 2 == 1
    half = nnz_per_row // 2
    pred = np.arange(nnz_per_row - half, nnz_per_row + 1)
    main_rows = N - 2 * (nnz_per_row - half)
    nnz_arr = np.concatenate((pred, np.ones(main_rows) * nnz_per_row, pred[::-1]))
    indptr = np.zeros(N + 1, dtype=np.int64)
    indptr[1:] = np.cumsum(nnz_arr)
    cols = np.tile(np.arange(-half, nnz_per_row - half), (N,)) + np.repeat(np.arange(N), nnz_per_row)
    data = np.ones(N * nnz_per_row, dtype=dtype) if ones else (np.arange(N * nnz_per_row) / N).astype(dtype)
    mask = (cols >= 0) & (cols < N)
    return data[mask], cols[mask].astype(np.int64), indptr


def random_csr_fixed(n, m, k, seed=1234, dtype=np.float64):
    """n x m, exactly k nnz per row: column j of a row is drawn uniformly from the j-th of k
    equal strata of [0, m) → distinct, sorted, spread over the whole x (C2's 'random CSR
    ~50 nnz/row'); values standard normal."""
    rng = np.random.default_rng(seed)
    stride = m // k
    cols = (np.arange(k, dtype=np.int64) * stride)[None, :] + rng.integers(0, stride, size=(n, k))
    data = rng.standard_normal(n * k).astype(dtype)
    indptr = np.arange(n + 1, dtype=np.int64) * k
    return data, cols.reshape(-1), indptr


def powerlaw_csr(n, m, max_row=10_000, alpha=2.0, seed=7, mean_hint=None, dtype=np.float64):
    """Power-law row degrees (Zipf alpha) clipped to [1, max_row], at least one row at
    max_row; uniform columns (C5)."""
    rng = np.random.default_rng(seed)
    deg = np.minimum(rng.zipf(alpha, size=n), max_row).astype(np.int64)
    deg[rng.integers(0, n)] = max_row
    deg = np.minimum(deg, m)
    indptr = np.zeros(n + 1, dtype=np.int64)
    np.cumsum(deg, out=indptr[1:])
    nnz = int(indptr[-1])
    cols = rng.integers(0, m, size=nnz).astype(np.int64)
    data = rng.standard_normal(nnz).astype(dtype)
    return data, cols, indptr


def rmat_csr(scale, edge_factor=16, a=0.57, b=0.19, c=0.19, seed=42, dtype=np.float64):
    """R-MAT graph (not in the reference — BASELINE config 4): 2^scale vertices,
    edge_factor*2^scale edges, duplicates summed (values 1.0)."""
    rng = np.random.default_rng(seed)
    n = 1 << scale
    ne = edge_factor * n
    rows = np.zeros(ne, dtype=np.int64)
    cols = np.zeros(ne, dtype=np.int64)
    for _ in range(scale):
        r = rng.random(ne)
        rbit = r >= a + b
        cbit = ((r >= a) & (r < a + b)) | (r >= a + b + c)
        rows = (rows << 1) | rbit
        cols = (cols << 1) | cbit
    M = sp.coo_array((np.ones(ne, dtype=dtype), (rows, cols)), shape=(n, n)).tocsr()
    M.sum_duplicates()
    M.sort_indices()
    return M


def simple_system(N, M, seed=0, tol=0.5):
    """rand(N,M) thresholded to ~50% density + rand(M) vector — the reference's
    simple_system_gen (tests/integration/utils/sample.py:48-55) with a numpy RNG."""
    rng = np.random.default_rng(seed)
    a = rng.random((N, M))
    x = rng.random(M)
    a = np.where(a < tol, a, 0)
    return a, x
