// b2s_convert.cu — small CSR helpers on the edges of the hot path.
#include "b2s_common.cuh"

namespace b2s {

// GetCSRDiagonal (reference get_diagonal.cu:25-44, get_diagonal.cc:32-41):
// diag[i] = vals[j] for the LAST j in row i with crd[j] == i, else 0.
// One 8-lane group per row so the col reads of a row are coalesced.
template <typename V, typename I>
__global__ void diagonal_kernel(int64_t nrows, const int64_t* __restrict__ indptr,
                                const I* __restrict__ cols, const V* __restrict__ vals,
                                V* __restrict__ diag) {
  constexpr int L = 8;
  int64_t g = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) / L;
  int gl = threadIdx.x & (L - 1);
  int64_t total = ((int64_t)gridDim.x * blockDim.x) / L;
  int64_t nr_round = ceil_div(nrows, total) * total;  // warp-uniform trip count
  for (int64_t r = g; r < nr_round; r += total) {
    int64_t best = -1;
    if (r < nrows) {
      int64_t lo = indptr[r], hi = indptr[r + 1];
      for (int64_t p = lo + gl; p < hi; p += L)
        if ((int64_t)cols[p] == r) best = p;
    }
#pragma unroll
    for (int o = L >> 1; o > 0; o >>= 1) {
      int64_t other = __shfl_xor_sync(0xffffffffu, best, o);
      best = other > best ? other : best;
    }
    if (r < nrows && gl == 0) diag[r] = best >= 0 ? vals[best] : zero_of<V>();
  }
}

// ExpandPosToCoordinates (reference pos_to_coordinates_template.inl:46-112, a thrust
// fill/scatter/scan/gather pipeline) as one kernel: an 8-lane group per row writes its row id.
__global__ void expand_rows_kernel(int64_t nrows, const int64_t* __restrict__ indptr,
                                   int64_t* __restrict__ rows_out) {
  constexpr int L = 8;
  int64_t g = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) / L;
  int gl = threadIdx.x & (L - 1);
  int64_t total = ((int64_t)gridDim.x * blockDim.x) / L;
  for (int64_t r = g; r < nrows; r += total) {
    int64_t lo = indptr[r], hi = indptr[r + 1];
    for (int64_t p = lo + gl; p < hi; p += L) rows_out[p] = r;
  }
}

template <typename S, typename D>
__global__ void cast_kernel(int64_t n, const S* __restrict__ src, D* __restrict__ dst) {
  int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) dst[i] = (D)src[i];
}

// CSRToDense (reference csr_to_dense.cu:25-47): thread-per-entry scatter into a zeroed
// row-major matrix.  Duplicate (row, col) entries: the reference's loop keeps the last one
// written (A_vals[...] = on a sequential row loop); here the highest position wins as well
// because a row is handled by ONE thread in order.
template <typename V, typename I>
__global__ void to_dense_kernel(int64_t nrows, int64_t ncols, const int64_t* __restrict__ indptr,
                                const I* __restrict__ cols, const V* __restrict__ vals,
                                V* __restrict__ out) {
  int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= nrows) return;
  for (int64_t p = indptr[r]; p < indptr[r + 1]; ++p) out[r * ncols + (int64_t)cols[p]] = vals[p];
}

static inline int64_t grid_for(int64_t work, int threads) {
  int64_t b = ceil_div(work, threads);
  int64_t cap = (int64_t)kNumSMs * 16;
  if (b > cap) b = cap;
  return b < 1 ? 1 : b;
}

}  // namespace b2s

using namespace b2s;

extern "C" int b2s_csr_diagonal(b2s_dtype vt, b2s_itype it, int64_t nrows, const int64_t* indptr,
                                const void* indices, const void* data, void* diag,
                                b2s_stream_t stream) {
  B2S_REQUIRE(nrows >= 0, "negative nrows");
  if (nrows == 0) return B2S_OK;
  B2S_REQUIRE(indptr && diag, "null pointer");
  cudaStream_t st = (cudaStream_t)stream;
  B2S_DISPATCH_VT(vt, V, B2S_DISPATCH_IT(it, I, {
    diagonal_kernel<V, I><<<(unsigned)grid_for(nrows * 8, 256), 256, 0, st>>>(
        nrows, indptr, (const I*)indices, (const V*)data, (V*)diag);
    B2S_CHECK_LAUNCH();
  }));
  return B2S_OK;
}

extern "C" int b2s_csr_expand_rows(int64_t nrows, int64_t nnz, const int64_t* indptr,
                                   int64_t* rows_out, b2s_stream_t stream) {
  B2S_REQUIRE(nrows >= 0 && nnz >= 0, "negative size");
  if (nrows == 0 || nnz == 0) return B2S_OK;
  B2S_REQUIRE(indptr && rows_out, "null pointer");
  expand_rows_kernel<<<(unsigned)grid_for(nrows * 8, 256), 256, 0, (cudaStream_t)stream>>>(nrows, indptr,
                                                                                           rows_out);
  B2S_CHECK_LAUNCH();
  return B2S_OK;
}

extern "C" int b2s_cast_i64_to_i32(int64_t n, const int64_t* src, int32_t* dst, b2s_stream_t stream) {
  B2S_REQUIRE(n >= 0, "negative n");
  if (n == 0) return B2S_OK;
  B2S_REQUIRE(src && dst, "null pointer");
  cast_kernel<int64_t, int32_t><<<(unsigned)grid_for(n, 256), 256, 0, (cudaStream_t)stream>>>(n, src, dst);
  B2S_CHECK_LAUNCH();
  return B2S_OK;
}

extern "C" int b2s_cast_i32_to_i64(int64_t n, const int32_t* src, int64_t* dst, b2s_stream_t stream) {
  B2S_REQUIRE(n >= 0, "negative n");
  if (n == 0) return B2S_OK;
  B2S_REQUIRE(src && dst, "null pointer");
  cast_kernel<int32_t, int64_t><<<(unsigned)grid_for(n, 256), 256, 0, (cudaStream_t)stream>>>(n, src, dst);
  B2S_CHECK_LAUNCH();
  return B2S_OK;
}

extern "C" int b2s_csr_to_dense(b2s_dtype vt, b2s_itype it, int64_t nrows, int64_t ncols,
                                const int64_t* indptr, const void* indices, const void* data,
                                void* out, b2s_stream_t stream) {
  B2S_REQUIRE(nrows >= 0 && ncols >= 0, "negative size");
  if (nrows == 0 || ncols == 0) return B2S_OK;
  B2S_REQUIRE(indptr && out, "null pointer");
  cudaStream_t st = (cudaStream_t)stream;
  B2S_CUDA_TRY(cudaMemsetAsync(out, 0, (size_t)nrows * (size_t)ncols * dtype_size(vt), st));
  B2S_DISPATCH_VT(vt, V, B2S_DISPATCH_IT(it, I, {
    to_dense_kernel<V, I><<<(unsigned)ceil_div(nrows, 128), 128, 0, st>>>(
        nrows, ncols, indptr, (const I*)indices, (const V*)data, (V*)out);
    B2S_CHECK_LAUNCH();
  }));
  return B2S_OK;
}
