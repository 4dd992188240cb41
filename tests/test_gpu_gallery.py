"""Device-side constructors and legate_sparse.random (g1):
dense→CSR (reference dense_to_csr.cu:25-43,128-149), DIA→CSR (dia.py:159-190), CSR→dense
(csr_to_dense.cu:25-47) and the counter-based random generator, each against the host statement
of the same algorithm (oracle / numpy path) bit for bit, and against scipy."""
import numpy as np
import pytest
import scipy.sparse as sp
import torch

import legate_sparse as sparse
from oracle import oracle

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("m,n,density,seed", [(1000, 800, 0.0154, 7), (64, 5000, 0.3, 1), (5000, 64, 0.01, 2),
                                              (3, 3, 1.0, 9), (257, 1031, 0.5, 123456789012345)])
def test_random_matches_host_twin_bit_for_bit(m, n, density, seed):
    A = sparse.random(m, n, density=density, rng=seed, dtype=np.float64)
    k = int(round(density * m * n))
    p, c, v = oracle.random_csr(m, n, k, seed)
    assert A.shape == (m, n) and A.nnz == k and A.dtype == np.float64
    assert np.array_equal(A.indptr, p) and np.array_equal(A.indices, c)
    assert np.array_equal(A.data, v)                      # same fma on both sides: bit-exact
    assert A.has_sorted_indices() and A.has_canonical_format()
    # a row block generated on its own equals the slice of the whole
    r0, r1 = m // 3, m - m // 4
    B = sparse.random(m, n, density=density, rng=seed, row_block=(r0, r1))
    blk = B._block()
    assert (blk.r0, blk.r1) == (r0, r1)
    assert np.array_equal(blk.indices.cpu().numpy(), c[p[r0]:p[r1]])
    assert np.array_equal(blk.data.cpu().numpy(), v[p[r0]:p[r1]])


def test_random_scipy_parity_properties():
    """same contract as scipy.sparse.random: nnz = round(density*m*n), values uniform in [0,1),
    positions uniform; deterministic per seed; data_rvs honoured; dtype honoured"""
    m, n, density = 2000, 3000, 0.01
    S = sp.random(m, n, density=density, format="csr", random_state=3)
    A = sparse.random(m, n, density=density, format="csr", rng=3)
    assert A.nnz == S.nnz and A.shape == S.shape and A.dtype == S.dtype
    d = A.data
    assert 0.0 <= d.min() and d.max() < 1.0 and abs(d.mean() - 0.5) < 0.01 and abs(d.var() - 1 / 12) < 0.005
    col_hist = np.bincount(A.indices, minlength=n)
    assert abs(col_hist.mean() - S.nnz / n) < 1e-9 and col_hist.std() < 3.0 * np.sqrt(S.nnz / n)
    rows = np.diff(A.indptr)
    assert rows.max() - rows.min() <= 1                      # spread as evenly as possible
    A2 = sparse.random(m, n, density=density, rng=3)
    assert np.array_equal(A2.indices, A.indices) and np.array_equal(A2.data, A.data)
    A3 = sparse.random(m, n, density=density, rng=4)
    assert not np.array_equal(A3.indices, A.indices)
    g = sparse.random(m, n, density=density, rng=np.random.default_rng(5))
    assert g.nnz == S.nnz
    ones = sparse.random(m, n, density=density, rng=3, data_rvs=lambda k: np.ones(k))
    assert np.array_equal(ones.indices, A.indices) and np.all(ones.data == 1.0)
    for dt in (np.float32, np.complex64, np.complex128):
        B = sparse.random(300, 200, density=0.05, rng=1, dtype=dt)
        assert B.dtype == np.dtype(dt) and B.nnz == 3000
        if np.dtype(dt).kind == "c":
            assert np.abs(B.data.imag).max() > 0
    with pytest.raises(NotImplementedError):
        sparse.random(10, 10, format="coo")
    with pytest.raises(ValueError):
        sparse.random(10, 10, density=1.5)
    x = np.random.default_rng(0).standard_normal(n)
    assert np.allclose(A @ x, sp.csr_array((A.data, A.indices, A.indptr), shape=A.shape) @ x, rtol=1e-12)


@pytest.mark.parametrize("dtype", [np.float32, np.float64, np.complex64, np.complex128])
@pytest.mark.parametrize("shape", [(1, 1), (37, 129), (300, 33), (64, 64)])
def test_dense_to_csr_on_device(dtype, shape):
    rng = np.random.default_rng(3)
    D = rng.standard_normal(shape).astype(dtype)
    if np.dtype(dtype).kind == "c":
        D = D + 1j * rng.standard_normal(shape).astype(D.real.dtype)
    D[rng.random(shape) < 0.7] = 0
    if shape[0] > 2:
        D[1, :] = 0                      # an empty row
    D.flat[0] = -0.0                     # negative zero is a zero (`!= 0` test)
    Dt = torch.from_numpy(D).cuda()
    A = sparse.csr_array(Dt)
    assert A._h_data is None and A._g_data is not None and A._g_data.is_cuda     # no host round trip
    H = sparse.csr_array(D)              # host numpy statement of dense_to_csr.cc:32-64
    assert A.dtype == np.dtype(dtype) and A.shape == shape
    assert np.array_equal(A.indptr, H.indptr) and np.array_equal(A.indices, H.indices)
    assert np.array_equal(A.data, H.data)
    o = oracle.dense_to_csr(D)
    assert np.array_equal(A.indptr, o[0]) and np.array_equal(A.indices, o[1])
    # and back: CSR → dense on the device
    assert np.array_equal(A.todense(), D + 0.0)
    out = torch.full(shape, 7, dtype=Dt.dtype, device="cuda")
    A.todense(out=out)
    assert torch.equal(out, torch.from_numpy(D + 0.0).cuda())
    with pytest.raises(ValueError):
        A.todense(out=np.zeros(shape, dtype=np.int32))


def test_dense_nan_is_a_stored_entry():
    D = np.zeros((4, 5))
    D[2, 3] = np.nan
    A = sparse.csr_array(torch.from_numpy(D).cuda())
    assert A.nnz == 1 and A.indices[0] == 3 and np.isnan(A.data[0])


@pytest.mark.parametrize("dtype", [np.float32, np.float64, np.complex128])
def test_dia_to_csr_on_device_equals_host_statement(dtype):
    rng = np.random.default_rng(8)
    cases = [
        ([-1, 0, 1], (50, 50)), ([1, 0, -1], (50, 50)), ([-7, -2, 3, 11], (40, 65)), ([0], (9, 9)),
        ([-30, 5], (33, 12)), ([2, -2, 0, 40], (64, 41)), ([-4096, -1, 0, 1, 4096], (4096 * 3, 4096 * 3)),
    ]
    for offs, shape in cases:
        width = min(shape[1], max(shape))
        data = rng.standard_normal((len(offs), width)).astype(dtype)
        data[rng.random(data.shape) < 0.2] = 0                 # explicit zeros are dropped
        Dm = sparse.dia_array((data, np.array(offs)), shape=shape, dtype=dtype)
        dev = Dm.tocsr()
        host = Dm._tocsr_host()
        assert dev._g_data is not None and dev._h_data is None      # built on the device
        assert np.array_equal(dev.indptr, host.indptr), (offs, shape)
        assert np.array_equal(dev.indices, host.indices), (offs, shape)
        assert np.array_equal(dev.data, host.data), (offs, shape)
        S = sp.dia_array((data, offs), shape=shape).tocsr()
        S.eliminate_zeros()
        S.sort_indices()
        assert np.array_equal(dev.indptr, S.indptr) and np.array_equal(dev.indices, S.indices)
        assert np.array_equal(dev.data, S.data)


def test_diags_poisson_config1_built_on_device():
    n = 1000
    main = np.full(n * n, 4.0)
    off1 = np.full(n * n - 1, -1.0)
    off1[np.arange(1, n * n) % n == 0] = 0
    offn = np.full(n * n - n, -1.0)
    A = sparse.diags([main, off1, off1, offn, offn], [0, -1, 1, -n, n], shape=(n * n, n * n), format="csr",
                     dtype=np.float64)
    S = sp.diags([main, off1, off1, offn, offn], [0, -1, 1, -n, n], shape=(n * n, n * n), format="csr",
                 dtype=np.float64)
    S.eliminate_zeros()
    assert A._g_data is not None and A._h_data is None
    assert np.array_equal(A.indptr, S.indptr) and np.array_equal(A.indices, S.indices)   # bit-exact structure
    assert np.array_equal(A.data, S.data)
