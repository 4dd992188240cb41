// b2s_gallery.cu — device-side constructors on the edges of the hot path (SURVEY §8 row f4 + the
// `random` generator north_star names):
//   * dense → CSR, two passes with a `!= 0` test   (reference src/sparse/array/conv/dense_to_csr.cu:25-43
//     count, :128-149 fill; CPU loops dense_to_csr.cc:32-40,55-64)
//   * DIA → CSR                                    (reference legate_sparse/dia.py:159-190, cupynumeric ops)
//   * counter-based random CSR generator           (legate_sparse.random; the reference has none — its
//     tests draw from cupynumeric's RNG and densify, utils/sample.py:21-45)
// All are stream-ordered; the count passes write per-row counts that the caller scans with
// b2s_scan_i64 (one native scan, no host round trip except the final nnz).
#include "b2s_common.cuh"

namespace b2s {

// ------------------------------------------------------------------ dense → CSR
// One warp per row; lanes stride over the columns (coalesced), ballots keep column order.
template <typename V>
__global__ void __launch_bounds__(256)
dense_count_kernel(int64_t nrows, int64_t ncols, int64_t ld, const V* __restrict__ dense,
                   int64_t* __restrict__ row_nnz /* [nrows] */) {
  const int lane = threadIdx.x & 31;
  const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t r = warp; r < nrows; r += nwarps) {
    const V* row = dense + r * ld;
    int64_t cnt = 0;
    for (int64_t c0 = 0; c0 < ncols; c0 += 32) {
      const int64_t c = c0 + lane;
      const bool nz = c < ncols && !vis_zero(row[c]);
      cnt += __popc(__ballot_sync(0xffffffffu, nz));
    }
    if (lane == 0) row_nnz[r] = cnt;
  }
}

template <typename V, typename I>
__global__ void __launch_bounds__(256)
dense_fill_kernel(int64_t nrows, int64_t ncols, int64_t ld, const V* __restrict__ dense,
                  const int64_t* __restrict__ indptr, I* __restrict__ cols, V* __restrict__ vals) {
  const int lane = threadIdx.x & 31;
  const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t r = warp; r < nrows; r += nwarps) {
    const V* row = dense + r * ld;
    int64_t out = indptr[r];
    for (int64_t c0 = 0; c0 < ncols; c0 += 32) {
      const int64_t c = c0 + lane;
      V v = zero_of<V>();
      bool nz = false;
      if (c < ncols) { v = row[c]; nz = !vis_zero(v); }
      const unsigned m = __ballot_sync(0xffffffffu, nz);
      if (nz) {
        const int64_t dst = out + __popc(m & ((1u << lane) - 1u));
        cols[dst] = (I)c;
        vals[dst] = v;
      }
      out += __popc(m);
    }
  }
}

// ------------------------------------------------------------------ DIA → CSR
// data[d][j] (leading dimension ld) is A[j - offsets[d], j].  One thread per row walks the
// diagonals in the given order (`order` = positions of the offsets sorted ascending, so columns come
// out ascending: the canonical sorted CSR scipy builds); explicit zeros are dropped (dia.py:171).
template <typename V>
__global__ void dia_count_kernel(int64_t nrows, int64_t ncols, int ndiag, int64_t width, int64_t ld,
                                 const V* __restrict__ data, const int64_t* __restrict__ offsets,
                                 const int* __restrict__ order, int64_t* __restrict__ row_nnz) {
  const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= nrows) return;
  int64_t cnt = 0;
  for (int q = 0; q < ndiag; ++q) {
    const int d = order[q];
    const int64_t j = r + offsets[d];
    if (j >= 0 && j < ncols && j < width && !vis_zero(data[(int64_t)d * ld + j])) ++cnt;
  }
  row_nnz[r] = cnt;
}

template <typename V, typename I>
__global__ void dia_fill_kernel(int64_t nrows, int64_t ncols, int ndiag, int64_t width, int64_t ld,
                                const V* __restrict__ data, const int64_t* __restrict__ offsets,
                                const int* __restrict__ order, const int64_t* __restrict__ indptr,
                                I* __restrict__ cols, V* __restrict__ vals) {
  const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= nrows) return;
  int64_t out = indptr[r];
  for (int q = 0; q < ndiag; ++q) {
    const int d = order[q];
    const int64_t j = r + offsets[d];
    if (j >= 0 && j < ncols && j < width) {
      const V v = data[(int64_t)d * ld + j];
      if (!vis_zero(v)) { cols[out] = (I)j; vals[out] = v; ++out; }
    }
  }
}

// ------------------------------------------------------------------ random CSR (counter-based)
// Exactly nnz_total stored entries in an m x n matrix: row i holds k_i = q + [((i + shift) mod m) < rem]
// entries (q = nnz_total / m, rem = nnz_total % m, shift = mix64(seed) mod m); its j-th entry lies in
// the j-th of k_i equal strata of [0, n): columns are distinct, sorted and marginally uniform.
// Entry (i, j) depends only on (seed, i, j) → any row block can be generated independently by any
// rank, and oracle/ref_kernels.c:ref_random_csr reproduces it bit for bit on the host.
__host__ __device__ __forceinline__ uint64_t mix64(uint64_t t) {
  t = (t ^ (t >> 30)) * 0xBF58476D1CE4E5B9ull;
  t = (t ^ (t >> 27)) * 0x94D049BB133111EBull;
  return t ^ (t >> 31);
}
struct RandomLayout {
  int64_t m, q, rem, shift;
  __host__ __device__ int64_t extras_before(int64_t a) const {   // #u in [0,a) with u mod m < rem
    return (a / m) * rem + ((a % m) < rem ? (a % m) : rem);
  }
  __host__ __device__ int64_t row_start(int64_t i) const {       // global position of row i's first entry
    return i * q + extras_before(i + shift) - extras_before(shift);
  }
  __host__ __device__ int64_t row_len(int64_t i) const { return q + (((i + shift) % m) < rem ? 1 : 0); }
};
static RandomLayout random_layout(int64_t m, int64_t nnz_total, uint64_t seed) {
  RandomLayout L;
  L.m = m; L.q = nnz_total / m; L.rem = nnz_total % m;
  L.shift = (int64_t)(mix64(seed) % (uint64_t)m);
  return L;
}

__global__ void random_rowptr_kernel(RandomLayout L, int64_t r0, int64_t nloc, int64_t* __restrict__ indptr) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i > nloc) return;
  indptr[i] = L.row_start(r0 + i) - L.row_start(r0);
}

__device__ __forceinline__ double u01(uint64_t h) { return (double)(h >> 11) * (1.0 / 9007199254740992.0); }
__device__ __forceinline__ void rand_value(uint64_t h, double lo, double hi, float* out) { *out = (float)fma(hi - lo, u01(h), lo); }
__device__ __forceinline__ void rand_value(uint64_t h, double lo, double hi, double* out) { *out = fma(hi - lo, u01(h), lo); }
__device__ __forceinline__ void rand_value(uint64_t h, double lo, double hi, c64* out) {
  out->re = (float)fma(hi - lo, u01(h), lo);
  out->im = (float)fma(hi - lo, u01(mix64(h + 0x9E3779B97F4A7C15ull)), lo);
}
__device__ __forceinline__ void rand_value(uint64_t h, double lo, double hi, c128* out) {
  out->re = fma(hi - lo, u01(h), lo);
  out->im = fma(hi - lo, u01(mix64(h + 0x9E3779B97F4A7C15ull)), lo);
}

// one warp per row, lanes stride over the row's entries (coalesced stores)
template <typename V, typename I>
__global__ void __launch_bounds__(256)
random_fill_kernel(RandomLayout L, int64_t n, uint64_t seed, int64_t r0, int64_t nloc, double lo, double hi,
                   I* __restrict__ cols, V* __restrict__ vals) {
  const int lane = threadIdx.x & 31;
  const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  const int64_t base0 = L.row_start(r0);
  for (int64_t rl = warp; rl < nloc; rl += nwarps) {
    const int64_t i = r0 + rl;
    const int64_t k = L.row_len(i);
    const int64_t base = L.row_start(i) - base0;
    const uint64_t rowkey = mix64(seed + 0x9E3779B97F4A7C15ull * (uint64_t)(i + 1));
    for (int64_t j = lane; j < k; j += 32) {
      const uint64_t s0 = (uint64_t)(((unsigned __int128)(uint64_t)j * (uint64_t)n) / (uint64_t)k);
      const uint64_t s1 = (uint64_t)(((unsigned __int128)(uint64_t)(j + 1) * (uint64_t)n) / (uint64_t)k);
      const uint64_t h = mix64(rowkey + (uint64_t)j);
      cols[base + j] = (I)(s0 + h % (s1 - s0));
      rand_value(mix64(h ^ 0x632BE59BD9B4E019ull), lo, hi, &vals[base + j]);
    }
  }
}

static inline int64_t warp_grid(int64_t rows) {
  int64_t b = ceil_div(rows, 8);
  const int64_t cap = (int64_t)kNumSMs * 32;
  if (b > cap) b = cap;
  return b < 1 ? 1 : b;
}

}  // namespace b2s

using namespace b2s;

// =================================================================== C ABI
extern "C" int64_t b2s_scan_workspace_bytes(int64_t n) {
  if (n < 0) return -1;
  return (ceil_div(n > 0 ? n : 1, 1024) + 1) * 8 + 256;
}

// indptr[0] = 0, indptr[i+1] = sum(counts[0..i]) — `indptr` holds the per-row counts in
// indptr[1..n] on entry (what the count passes below write when given indptr + 1).
extern "C" int b2s_scan_i64(int64_t n, int64_t* indptr, void* workspace, int64_t workspace_bytes,
                            b2s_stream_t stream) {
  B2S_REQUIRE(n >= 0, "negative n");
  B2S_REQUIRE(indptr != nullptr, "indptr is null");
  if (n == 0) { B2S_CUDA_TRY(cudaMemsetAsync(indptr, 0, 8, (cudaStream_t)stream)); return B2S_OK; }
  B2S_REQUIRE(workspace != nullptr && workspace_bytes >= b2s_scan_workspace_bytes(n), "scan workspace too small");
  int64_t* ws = reinterpret_cast<int64_t*>(((uintptr_t)workspace + 63) & ~(uintptr_t)63);
  return scan_inclusive_i64(n, indptr + 1, indptr, ws, (cudaStream_t)stream);
}

extern "C" int b2s_dense_to_csr_count(b2s_dtype vt, int64_t nrows, int64_t ncols, int64_t ld,
                                      const void* dense, int64_t* row_nnz, b2s_stream_t stream) {
  B2S_REQUIRE(nrows >= 0 && ncols >= 0 && ld >= ncols, "bad dense shape");
  if (nrows == 0) return B2S_OK;
  B2S_REQUIRE(row_nnz != nullptr && (ncols == 0 || dense != nullptr), "null pointer");
  cudaStream_t st = (cudaStream_t)stream;
  B2S_DISPATCH_VT(vt, V, {
    dense_count_kernel<V><<<(unsigned)warp_grid(nrows), 256, 0, st>>>(nrows, ncols, ld, (const V*)dense, row_nnz);
    B2S_CHECK_LAUNCH();
  });
  return B2S_OK;
}

extern "C" int b2s_dense_to_csr_fill(b2s_dtype vt, b2s_itype it, int64_t nrows, int64_t ncols, int64_t ld,
                                     const void* dense, const int64_t* indptr, void* indices, void* data,
                                     b2s_stream_t stream) {
  B2S_REQUIRE(nrows >= 0 && ncols >= 0 && ld >= ncols, "bad dense shape");
  if (nrows == 0 || ncols == 0) return B2S_OK;
  B2S_REQUIRE(dense && indptr, "null pointer");
  cudaStream_t st = (cudaStream_t)stream;
  B2S_DISPATCH_VT(vt, V, B2S_DISPATCH_IT(it, I, {
    dense_fill_kernel<V, I><<<(unsigned)warp_grid(nrows), 256, 0, st>>>(nrows, ncols, ld, (const V*)dense, indptr,
                                                                       (I*)indices, (V*)data);
    B2S_CHECK_LAUNCH();
  }));
  return B2S_OK;
}

extern "C" int b2s_dia_to_csr_count(b2s_dtype vt, int64_t nrows, int64_t ncols, int ndiag, int64_t width,
                                    int64_t ld, const void* data, const int64_t* offsets, const int* order,
                                    int64_t* row_nnz, b2s_stream_t stream) {
  B2S_REQUIRE(nrows >= 0 && ncols >= 0 && ndiag >= 0 && width >= 0 && ld >= width, "bad DIA shape");
  if (nrows == 0) return B2S_OK;
  B2S_REQUIRE(row_nnz != nullptr && (ndiag == 0 || (data && offsets && order)), "null pointer");
  cudaStream_t st = (cudaStream_t)stream;
  B2S_DISPATCH_VT(vt, V, {
    dia_count_kernel<V><<<(unsigned)ceil_div(nrows, 256), 256, 0, st>>>(nrows, ncols, ndiag, width, ld, (const V*)data,
                                                                       offsets, order, row_nnz);
    B2S_CHECK_LAUNCH();
  });
  return B2S_OK;
}

extern "C" int b2s_dia_to_csr_fill(b2s_dtype vt, b2s_itype it, int64_t nrows, int64_t ncols, int ndiag,
                                   int64_t width, int64_t ld, const void* data, const int64_t* offsets,
                                   const int* order, const int64_t* indptr, void* indices, void* out_data,
                                   b2s_stream_t stream) {
  B2S_REQUIRE(nrows >= 0 && ncols >= 0 && ndiag >= 0 && width >= 0 && ld >= width, "bad DIA shape");
  if (nrows == 0 || ndiag == 0) return B2S_OK;
  B2S_REQUIRE(data && offsets && order && indptr, "null pointer");
  cudaStream_t st = (cudaStream_t)stream;
  B2S_DISPATCH_VT(vt, V, B2S_DISPATCH_IT(it, I, {
    dia_fill_kernel<V, I><<<(unsigned)ceil_div(nrows, 256), 256, 0, st>>>(nrows, ncols, ndiag, width, ld, (const V*)data,
                                                                         offsets, order, indptr, (I*)indices,
                                                                         (V*)out_data);
    B2S_CHECK_LAUNCH();
  }));
  return B2S_OK;
}

extern "C" int b2s_random_csr_rowptr(int64_t m, int64_t nnz_total, uint64_t seed, int64_t r0, int64_t r1,
                                     int64_t* indptr_local, b2s_stream_t stream) {
  B2S_REQUIRE(m > 0 && nnz_total >= 0 && r0 >= 0 && r0 <= r1 && r1 <= m, "bad row range");
  B2S_REQUIRE(indptr_local != nullptr, "indptr is null");
  const RandomLayout L = random_layout(m, nnz_total, seed);
  const int64_t nloc = r1 - r0;
  random_rowptr_kernel<<<(unsigned)ceil_div(nloc + 1, 256), 256, 0, (cudaStream_t)stream>>>(L, r0, nloc, indptr_local);
  B2S_CHECK_LAUNCH();
  return B2S_OK;
}

extern "C" int64_t b2s_random_csr_block_nnz(int64_t m, int64_t nnz_total, uint64_t seed, int64_t r0, int64_t r1) {
  if (m <= 0 || nnz_total < 0 || r0 < 0 || r0 > r1 || r1 > m) return -1;
  const RandomLayout L = random_layout(m, nnz_total, seed);
  return L.row_start(r1) - L.row_start(r0);
}

extern "C" int b2s_random_csr_fill(b2s_dtype vt, b2s_itype it, int64_t m, int64_t n, int64_t nnz_total,
                                   uint64_t seed, int64_t r0, int64_t r1, double lo, double hi, void* indices,
                                   void* data, b2s_stream_t stream) {
  B2S_REQUIRE(m > 0 && n > 0 && nnz_total >= 0 && r0 >= 0 && r0 <= r1 && r1 <= m, "bad shape / row range");
  const RandomLayout L = random_layout(m, nnz_total, seed);
  B2S_REQUIRE(L.q + (L.rem > 0 ? 1 : 0) <= n, "more entries per row than columns (density > 1)");
  const int64_t nloc = r1 - r0;
  if (nloc == 0 || L.row_start(r1) == L.row_start(r0)) return B2S_OK;
  B2S_REQUIRE(indices && data, "null pointer");
  cudaStream_t st = (cudaStream_t)stream;
  B2S_DISPATCH_VT(vt, V, B2S_DISPATCH_IT(it, I, {
    random_fill_kernel<V, I><<<(unsigned)warp_grid(nloc), 256, 0, st>>>(L, n, seed, r0, nloc, lo, hi, (I*)indices,
                                                                       (V*)data);
    B2S_CHECK_LAUNCH();
  }));
  return B2S_OK;
}
