"""oracle/oracle.py — TEST INFRASTRUCTURE ONLY (never imported by the product).

CPU restatement of the reference's algorithms for the CSR hot path:
  * C task bodies (oracle/ref_kernels.c via ctypes): SpMV, OpenMP SpMV, two-pass Gustavson
    SpGEMM, AXPBY, diagonal, row expansion — each citing the reference file:line it follows;
  * numpy restatements of the Python-level algorithms: CG (legate_sparse/linalg.py:465-535),
    GMRES (linalg.py:540-668), dense→CSR (dense_to_csr.cc:32-64), COO→CSR (csr.py:198-219),
    DIA→CSR / diags (dia.py:152-190, gallery.py:136-195), MatrixMarket reader
    (mtx_to_coo.cc:50-137).

Parity status ("pinned"): checked in tests/test_oracle_golden.py against
  (1) the reference's own known-answer vectors (tests/golden/reference_known_answers.json,
      transcribed from its tests/README with file:line), and
  (2) tests/golden/refrun_*.npz — outputs of the reference's OWN Python code
      (linalg.cg / linalg.gmres / gallery.diags / dia_array.tocsr) executed in the build
      container by tests/golden/make_golden.py with cupynumeric→numpy and Legate-task shims, and
  (3) scipy.sparse (the oracle BASELINE.json names).
The reference's native tasks themselves cannot be built here (legate.h / private
legate.core.internal), so there is no oracle/_ref binary.
"""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def build(force: bool = False) -> str:
    path = os.path.join(_HERE, "liboracle.so")
    src = os.path.join(_HERE, "ref_kernels.c")
    if force or not os.path.exists(path) or os.path.getmtime(path) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "liboracle.so"], stdout=subprocess.DEVNULL)
    return path


def _lib():
    global _LIB
    if _LIB is None:
        _LIB = ctypes.CDLL(build())
        _LIB.ref_omp_max_threads.restype = ctypes.c_int
        _LIB.ref_spgemm_nnz.restype = ctypes.c_int
        _LIB.ref_spgemm_f64.restype = ctypes.c_int
    return _LIB


def _p(a):
    return a.ctypes.data_as(ctypes.c_void_p)


def _i64(a):
    return np.ascontiguousarray(a, dtype=np.int64)


def omp_threads() -> int:
    return int(_lib().ref_omp_max_threads())


def omp_set_threads(n: int) -> None:
    _lib().ref_omp_set_threads(ctypes.c_int(int(n)))


# ------------------------------------------------------------------ C task bodies
def spmv(indptr, indices, data, x, omp: bool = False):
    """y = A x, reference spmv.cc:36-43 (or spmv_omp.cc:36-44 with omp=True)."""
    indptr, n = _i64(indptr), len(indptr) - 1
    data = np.ascontiguousarray(data)
    if data.dtype == np.float64:
        x = np.ascontiguousarray(x, dtype=np.float64)
        y = np.empty(n, dtype=np.float64)
        if omp and np.asarray(indices).dtype == np.int32:
            idx = np.ascontiguousarray(indices)
            _lib().ref_spmv_omp_f64_i32(ctypes.c_int64(n), _p(indptr), _p(idx), _p(data), _p(x), _p(y))
            return y
        idx = _i64(indices)
        fn = _lib().ref_spmv_omp_f64 if omp else _lib().ref_spmv_f64
        fn(ctypes.c_int64(n), _p(indptr), _p(idx), _p(data), _p(x), _p(y))
        return y
    if data.dtype == np.float32:
        idx = _i64(indices)
        x = np.ascontiguousarray(x, dtype=np.float32)
        y = np.empty(n, dtype=np.float32)
        _lib().ref_spmv_f32(ctypes.c_int64(n), _p(indptr), _p(idx), _p(data), _p(x), _p(y))
        return y
    # complex: same loop in numpy (small cases only)
    idx = _i64(indices)
    y = np.zeros(n, dtype=np.result_type(data.dtype, np.asarray(x).dtype))
    for i in range(n):
        s = y.dtype.type(0)
        for jp in range(indptr[i], indptr[i + 1]):
            s += data[jp] * x[idx[jp]]
        y[i] = s
    return y


def random_csr(m, n, nnz_total, seed, r0=0, r1=None, lo=0.0, hi=1.0):
    """Host twin of the device generator behind legate_sparse.random (b2s_gallery.cu): rows
    [r0, r1) as (indptr_local, indices int64, data f64).  Arrays are first-touched by the OpenMP
    threads that fill them (what bench.py's CPU arm wants)."""
    r1 = m if r1 is None else r1
    q, rem = divmod(int(nnz_total), int(m))
    # block nnz from the closed form (same as b2s_random_csr_block_nnz)
    mask = (1 << 64) - 1

    def mix(t):
        t = ((t ^ (t >> 30)) * 0xBF58476D1CE4E5B9) & mask
        t = ((t ^ (t >> 27)) * 0x94D049BB133111EB) & mask
        return t ^ (t >> 31)

    shift = mix(int(seed) & mask) % int(m)

    def extras(a):
        return (a // m) * rem + min(a % m, rem)

    def start(i):
        return i * q + extras(i + shift) - extras(shift)

    nloc = r1 - r0
    nnz_loc = start(r1) - start(r0)
    indptr = np.empty(nloc + 1, dtype=np.int64)
    crd = np.empty(nnz_loc, dtype=np.int64)
    vals = np.empty(nnz_loc, dtype=np.float64)
    _lib().ref_random_csr_f64(ctypes.c_int64(m), ctypes.c_int64(n), ctypes.c_int64(nnz_total),
                              ctypes.c_uint64(int(seed) & mask), ctypes.c_int64(r0), ctypes.c_int64(r1),
                              ctypes.c_double(lo), ctypes.c_double(hi), _p(indptr), _p(crd), _p(vals))
    return indptr, crd, vals


def fill_uniform(n, seed):
    x = np.empty(int(n), dtype=np.float64)
    _lib().ref_fill_uniform_f64(ctypes.c_int64(n), ctypes.c_uint64(int(seed) & ((1 << 64) - 1)), _p(x))
    return x


def spgemm(a_ptr, a_idx, a_val, b_ptr, b_idx, b_val, ncolsB):
    """C = A B, reference two-pass CPU path: NNZ task (spgemm_csr_csr_csr.cc:62-87), cumsum
    (base.py:67-87), numeric task (:134-158).  Output columns in first-touch order."""
    a_ptr, a_idx, b_ptr, b_idx = _i64(a_ptr), _i64(a_idx), _i64(b_ptr), _i64(b_idx)
    a_val = np.ascontiguousarray(a_val, dtype=np.float64)
    b_val = np.ascontiguousarray(b_val, dtype=np.float64)
    n = len(a_ptr) - 1
    nnz_row = np.zeros(n, dtype=np.int64)
    rc = _lib().ref_spgemm_nnz(ctypes.c_int64(n), ctypes.c_int64(ncolsB), _p(a_ptr), _p(a_idx), _p(b_ptr),
                               _p(b_idx), _p(nnz_row))
    assert rc == 0
    c_ptr = np.zeros(n + 1, dtype=np.int64)
    np.cumsum(nnz_row, out=c_ptr[1:])
    c_idx = np.empty(int(c_ptr[-1]), dtype=np.int64)
    c_val = np.empty(int(c_ptr[-1]), dtype=np.float64)
    rc = _lib().ref_spgemm_f64(ctypes.c_int64(n), ctypes.c_int64(ncolsB), _p(a_ptr), _p(a_idx), _p(a_val),
                               _p(b_ptr), _p(b_idx), _p(b_val), _p(c_ptr), _p(c_idx), _p(c_val))
    assert rc == 0
    return c_ptr, c_idx, c_val


def axpby(y, x, a, b, isalpha=True, negate=False):
    """In-place AXPBY, reference axpby.cc:34-44."""
    y = np.ascontiguousarray(y, dtype=np.float64)
    x = np.ascontiguousarray(x, dtype=np.float64)
    a = np.ascontiguousarray(np.asarray(a, dtype=np.float64).reshape(-1)[:1])
    b = np.ascontiguousarray(np.asarray(b, dtype=np.float64).reshape(-1)[:1])
    _lib().ref_axpby_f64(ctypes.c_int64(y.shape[0]), _p(y), _p(x), _p(a), _p(b), int(bool(isalpha)),
                         int(bool(negate)))
    return y


def diagonal(indptr, indices, data):
    indptr, indices = _i64(indptr), _i64(indices)
    data = np.ascontiguousarray(data, dtype=np.float64)
    n = len(indptr) - 1
    d = np.empty(n, dtype=np.float64)
    _lib().ref_diagonal_f64(ctypes.c_int64(n), _p(indptr), _p(indices), _p(data), _p(d))
    return d


def expand_rows(indptr):
    indptr = _i64(indptr)
    n = len(indptr) - 1
    out = np.empty(int(indptr[-1]), dtype=np.int64)
    _lib().ref_expand_rows(ctypes.c_int64(n), _p(indptr), _p(out))
    return out


# ------------------------------------------------------------------ constructors (numpy)
def dense_to_csr(a):
    """dense → CSR: count `!= 0` per row, then fill in row-major order
    (reference dense_to_csr.cc:32-40 count, :55-64 fill)."""
    a = np.asarray(a)
    indptr = np.zeros(a.shape[0] + 1, dtype=np.int64)
    idx, val = [], []
    for i in range(a.shape[0]):
        for j in range(a.shape[1]):
            if a[i, j] != 0:
                idx.append(j)
                val.append(a[i, j])
        indptr[i + 1] = len(idx)
    return indptr, np.asarray(idx, dtype=np.int64), np.asarray(val, dtype=a.dtype)


def coo_to_csr(data, row, col, nrows):
    """COO → CSR: stable argsort by row, duplicates kept, columns in input order inside a row
    (reference csr.py:198-219)."""
    row = np.asarray(row)
    order = np.argsort(row, kind="stable")
    indptr = np.append(np.array([0]), np.cumsum(np.bincount(row, minlength=nrows))).astype(np.int64)
    return indptr, np.asarray(col)[order].astype(np.int64), np.asarray(data)[order]


def dia_to_csr(data, offsets, shape):
    """DIA → CSR exactly as the reference: transpose the DIA (dia.py:114-147), then the masked
    gather of dia.py:159-190 (drops explicit zeros)."""
    data = np.atleast_2d(np.asarray(data))
    offsets = np.atleast_1d(np.asarray(offsets))
    num_rows, num_cols = shape
    max_dim = max(shape)
    # --- transpose (dia.py:124-147)
    t_off = -offsets
    r = np.arange(len(t_off), dtype=np.int64)[:, None]
    c = np.arange(num_rows, dtype=np.int64) - (t_off % max_dim)[:, None]
    pad = max(0, max_dim - data.shape[1])
    t_data = np.hstack((data, np.zeros((data.shape[0], pad), dtype=data.dtype)))[r, c]
    # --- _tocsr_transposed on the transposed operand of shape (num_cols, num_rows)
    t_rows, t_cols = num_cols, num_rows
    _, offset_len = t_data.shape
    offset_inds = np.arange(offset_len)
    row = offset_inds - t_off[:, None]
    mask = row >= 0
    mask &= row < t_rows
    mask &= offset_inds < t_cols
    mask &= t_data != 0
    indptr = np.zeros(t_cols + 1, dtype=np.int64)
    indptr[1 : offset_len + 1] = np.cumsum(mask.sum(axis=0, dtype=np.int64)[:t_cols])
    if offset_len < t_cols:
        indptr[offset_len + 1 :] = indptr[offset_len]
    indices = row.T[mask.T].astype(np.int64, copy=False)
    vals = t_data.T[mask.T]
    return indptr, indices, vals


def diags_to_dia(diagonals, offsets, shape, dtype):
    """gallery.diags data-array construction (reference gallery.py:136-189)."""
    if np.isscalar(offsets):
        diagonals = [np.atleast_1d(diagonals)]
        offsets = [offsets]
    else:
        diagonals = list(map(np.atleast_1d, diagonals))
    if shape is None:
        m = len(diagonals[0]) + abs(int(offsets[0]))
        shape = (m, m)
    m, n = shape
    M = max([min(m + int(o), n - int(o)) + max(0, int(o)) for o in offsets])
    M = max(0, M)
    data_arr = np.zeros((len(offsets), M), dtype=dtype)
    K = min(m, n)
    for j, diagonal in enumerate(diagonals):
        offset = int(offsets[j])
        k = max(0, offset)
        length = min(m + offset, n - offset, K)
        data_arr[j, k : k + length] = diagonal[..., :length]
    return data_arr, np.atleast_1d(offsets), (m, n)


def mmread_coo(path):
    """MatrixMarket coordinate reader, reference mtx_to_coo.cc:50-137."""
    with open(path) as f:
        head = f.readline().split()
        assert head[0] == "%%MatrixMarket" and head[1] == "matrix" and head[2] == "coordinate"
        field, symmetry = head[3], head[4]
        assert field in ("real", "pattern", "integer") and symmetry in ("general", "symmetric")
        line = f.readline()
        while line.startswith("%"):
            line = f.readline()
        m, n, _ = (int(t) for t in line.split()[:3])
        rows, cols, vals = [], [], []
        for line in f:
            t = line.split()
            if not t:
                continue
            cx, cy = int(t[0]), int(t[1])
            v = 1.0 if field == "pattern" else (float(int(t[2])) if field == "integer" else float(t[2]))
            rows.append(cx - 1); cols.append(cy - 1); vals.append(v)
            if symmetry == "symmetric" and cx != cy:
                rows.append(cy - 1); cols.append(cx - 1); vals.append(v)
    return m, n, np.asarray(rows, np.int64), np.asarray(cols, np.int64), np.asarray(vals, np.float64)


# ------------------------------------------------------------------ solvers (numpy)
def _get_atol_rtol(b_norm, tol=None, atol=0.0, rtol=1e-5):
    rtol = float(tol) if tol is not None else rtol
    if atol is None:
        atol = rtol
    return max(float(atol), float(rtol) * float(b_norm)), rtol


def cg(matvec, b, x0=None, tol=None, maxiter=None, M=None, callback=None, atol=0.0, rtol=1e-5,
       conv_test_iters=25):
    """Reference CG recurrence and stopping cadence, linalg.py:465-535, on numpy arrays with
    the AXPBY task body from ref_kernels.c.  `matvec`/`M` are callables v -> A v."""
    b = np.asarray(b, dtype=np.float64)
    bnrm2 = np.linalg.norm(b)
    atol, _ = _get_atol_rtol(bnrm2, tol, atol, rtol)
    n = b.shape[0]
    if maxiter is None:
        maxiter = n * 10
    x = np.zeros(n) if x0 is None else np.array(x0, dtype=np.float64)
    p = np.zeros(n)
    r = b - matvec(x)
    iters = 0
    rho = np.zeros(1)
    while iters < maxiter:
        z = r.copy() if M is None else M(r)
        rho1 = rho
        rho = np.array([r.dot(z)])
        if iters == 0:
            p[:] = z
        else:
            p = axpby(p, z, rho, rho1, isalpha=False, negate=False)
        q = matvec(p)
        pq = np.array([p.dot(q)])
        x = axpby(x, p, rho, pq, isalpha=True, negate=False)
        r = axpby(r, q, rho, pq, isalpha=True, negate=True)
        iters += 1
        if callback is not None:
            callback(x)
        if (iters % conv_test_iters == 0 or iters == (maxiter - 1)) and np.linalg.norm(r) < atol:
            break
    return x, iters


def gmres(matvec, b, x0=None, tol=None, restart=None, maxiter=None, M=None, atol=0.0, rtol=1e-5):
    """Reference restarted GMRES (CGS Arnoldi + lstsq), linalg.py:592-668."""
    b = np.asarray(b, dtype=np.float64)
    n = b.shape[0]
    Mv = (lambda v: v.copy()) if M is None else M
    x = np.zeros(n) if x0 is None else np.array(x0, dtype=np.float64)
    bnrm2 = np.linalg.norm(b)
    atol, _ = _get_atol_rtol(bnrm2, tol, atol, rtol)
    if maxiter is None:
        maxiter = n * 10
    if restart is None:
        restart = 20
    restart = min(restart, n)
    V = np.empty((n, restart))
    H = np.zeros((restart + 1, restart))
    e = np.zeros((restart + 1,))
    iters = 0
    while True:
        mx = Mv(x)
        r = b - matvec(mx)
        r_norm = np.linalg.norm(r)
        if r_norm <= atol or iters >= maxiter:
            break
        v = r / r_norm
        V[:, 0] = v
        e[0] = r_norm
        for j in range(restart):
            z = Mv(v)
            u = matvec(z)
            h = V[:, : j + 1].conj().T @ u
            u = u - V[:, : j + 1] @ h
            H[: j + 1, j] = h
            H[j + 1, j] = np.linalg.norm(u)
            if j + 1 < restart:
                v = u / H[j + 1, j]
                V[:, j + 1] = v
        y = np.linalg.lstsq(H, e, rcond=None)[0]
        x = x + V @ y
        iters += restart
    info = 0
    if iters == maxiter and not (r_norm <= atol):
        info = iters
    return mx, info
