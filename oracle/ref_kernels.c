/*
 * oracle/ref_kernels.c — TEST INFRASTRUCTURE ONLY.
 *
 * CPU restatement of the reference's task bodies for the CSR hot path, in plain C,
 * following the reference loops statement by statement (same iteration order, same
 * accumulation order), but over scipy-layout `indptr` instead of the Rect<1> `pos`
 * store (pos[i] = {indptr[i], indptr[i+1]-1}, reference legate_sparse/csr.py:238-251).
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg
 * may load this file's library.  The product (legate-sparse_b200/) never does.
 *
 * Parity status: the reference's native tasks cannot be compiled here (every TU includes
 * legate.h; the pinned legate.core.internal 24.11.01 is private — SURVEY F13), so this
 * restatement is pinned against the reference's own known-answer vectors
 * (tests/golden/, see tests/test_oracle_golden.py) and against scipy.sparse, the oracle
 * BASELINE.json names.
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

/* ---- SpMV: reference src/sparse/array/csr/spmv.cc:36-43 (CPU variant) ------------- */
void ref_spmv_f64(int64_t nrows, const int64_t* indptr, const int64_t* crd, const double* vals,
                  const double* x, double* y)
{
  for (int64_t i = 0; i < nrows; i++) {
    double sum = 0.0;
    for (int64_t j_pos = indptr[i]; j_pos < indptr[i + 1]; j_pos++) {
      int64_t j = crd[j_pos];
      sum += vals[j_pos] * x[j];
    }
    y[i] = sum;
  }
}

void ref_spmv_f32(int64_t nrows, const int64_t* indptr, const int64_t* crd, const float* vals,
                  const float* x, float* y)
{
  for (int64_t i = 0; i < nrows; i++) {
    float sum = 0.0f;
    for (int64_t j_pos = indptr[i]; j_pos < indptr[i + 1]; j_pos++) {
      int64_t j = crd[j_pos];
      sum += vals[j_pos] * x[j];
    }
    y[i] = sum;
  }
}

/* ---- SpMV, OpenMP variant: reference spmv_omp.cc:36-44
 *      (#pragma omp parallel for schedule(monotonic : dynamic, 128)) -------------------- */
void ref_spmv_omp_f64(int64_t nrows, const int64_t* indptr, const int64_t* crd,
                      const double* vals, const double* x, double* y)
{
#pragma omp parallel for schedule(monotonic : dynamic, 128)
  for (int64_t i = 0; i < nrows; i++) {
    double sum = 0.0;
    for (int64_t j_pos = indptr[i]; j_pos < indptr[i + 1]; j_pos++) {
      int64_t j = crd[j_pos];
      sum += vals[j_pos] * x[j];
    }
    y[i] = sum;
  }
}

/* same loop with 32-bit column ids (scipy's own index width; dispatch.h:56-77 instantiates it) */
void ref_spmv_omp_f64_i32(int64_t nrows, const int64_t* indptr, const int32_t* crd,
                          const double* vals, const double* x, double* y)
{
#pragma omp parallel for schedule(monotonic : dynamic, 128)
  for (int64_t i = 0; i < nrows; i++) {
    double sum = 0.0;
    for (int64_t j_pos = indptr[i]; j_pos < indptr[i + 1]; j_pos++) {
      sum += vals[j_pos] * x[crd[j_pos]];
    }
    y[i] = sum;
  }
}

void ref_omp_set_threads(int n)
{
#ifdef _OPENMP
  if (n > 0) omp_set_num_threads(n);
#else
  (void)n;
#endif
}

int ref_omp_max_threads(void)
{
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}

/* ---- SpGEMM symbolic: reference spgemm_csr_csr_csr.cc:62-87 (SpGEMMCSRxCSRxCSRNNZ).
 *      C = A * B.  Dense `already_set` marker over the column span, first-touch index list. */
int ref_spgemm_nnz(int64_t nrowsA, int64_t ncolsB, const int64_t* a_ptr, const int64_t* a_crd,
                   const int64_t* b_ptr, const int64_t* b_crd, int64_t* nnz_per_row)
{
  int64_t* index_list = (int64_t*)malloc(sizeof(int64_t) * (size_t)(ncolsB > 0 ? ncolsB : 1));
  char* already_set   = (char*)calloc((size_t)(ncolsB > 0 ? ncolsB : 1), 1);
  if (!index_list || !already_set) { free(index_list); free(already_set); return 1; }
  for (int64_t i = 0; i < nrowsA; i++) {
    int64_t index_list_size = 0;
    for (int64_t kA = a_ptr[i]; kA < a_ptr[i + 1]; kA++) {
      int64_t k = a_crd[kA];
      for (int64_t jB = b_ptr[k]; jB < b_ptr[k + 1]; jB++) {
        int64_t j = b_crd[jB];
        if (!already_set[j]) {
          index_list[index_list_size] = j;
          already_set[j]              = 1;
          index_list_size++;
        }
      }
    }
    int64_t row_nnzs = 0;
    for (int64_t index_loc = 0; index_loc < index_list_size; index_loc++) {
      already_set[index_list[index_loc]] = 0;
      row_nnzs++;
    }
    nnz_per_row[i] = row_nnzs;
  }
  free(index_list);
  free(already_set);
  return 0;
}

/* ---- SpGEMM numeric: reference spgemm_csr_csr_csr.cc:134-158 (dense workspace Gustavson;
 *      output columns in FIRST-TOUCH order, values accumulated k-then-j). ---------------- */
int ref_spgemm_f64(int64_t nrowsA, int64_t ncolsB, const int64_t* a_ptr, const int64_t* a_crd,
                   const double* a_vals, const int64_t* b_ptr, const int64_t* b_crd,
                   const double* b_vals, const int64_t* c_ptr, int64_t* c_crd, double* c_vals)
{
  size_t span         = (size_t)(ncolsB > 0 ? ncolsB : 1);
  int64_t* index_list = (int64_t*)malloc(sizeof(int64_t) * span);
  char* already_set   = (char*)calloc(span, 1);
  double* workspace   = (double*)calloc(span, sizeof(double));
  if (!index_list || !already_set || !workspace) {
    free(index_list); free(already_set); free(workspace);
    return 1;
  }
  for (int64_t i = 0; i < nrowsA; i++) {
    int64_t index_list_size = 0;
    for (int64_t kA = a_ptr[i]; kA < a_ptr[i + 1]; kA++) {
      int64_t k = a_crd[kA];
      for (int64_t jB = b_ptr[k]; jB < b_ptr[k + 1]; jB++) {
        int64_t j = b_crd[jB];
        if (!already_set[j]) {
          index_list[index_list_size] = j;
          already_set[j]              = 1;
          index_list_size++;
        }
        workspace[j] += a_vals[kA] * b_vals[jB];
      }
    }
    int64_t pC = c_ptr[i];
    for (int64_t index_loc = 0; index_loc < index_list_size; index_loc++) {
      int64_t j      = index_list[index_loc];
      already_set[j] = 0;
      c_crd[pC]      = j;
      c_vals[pC]     = workspace[j];
      pC++;
      workspace[j] = 0.0;
    }
  }
  free(index_list);
  free(already_set);
  free(workspace);
  return 0;
}

/* ---- AXPBY: reference src/sparse/linalg/axpby.cc:34-44 --------------------------------
 *      val = a[0] / b[0]; NEGATE → val = -1 * val;
 *      IS_ALPHA: y = val*x + y   else: y = x + val*y                                      */
void ref_axpby_f64(int64_t n, double* y, const double* x, const double* a, const double* b,
                   int isalpha, int negate)
{
  double val = a[0] / b[0];
  if (negate) { val = (double)(-1) * val; }
  for (int64_t i = 0; i < n; i++) {
    if (isalpha) {
      y[i] = val * x[i] + y[i];
    } else {
      y[i] = x[i] + val * y[i];
    }
  }
}

/* ---- GetCSRDiagonal: reference src/sparse/array/csr/get_diagonal.cc:32-41 -------------- */
void ref_diagonal_f64(int64_t nrows, const int64_t* indptr, const int64_t* crd, const double* vals,
                      double* diag)
{
  for (int64_t i = 0; i < nrows; i++) {
    diag[i] = 0.0;
    for (int64_t j_pos = indptr[i]; j_pos < indptr[i + 1]; j_pos++) {
      if (crd[j_pos] == i) { diag[i] = vals[j_pos]; }
    }
  }
}

/* ---- ExpandPosToCoordinates: reference pos_to_coordinates_template.inl:46-112 (result) -- */
void ref_expand_rows(int64_t nrows, const int64_t* indptr, int64_t* rows_out)
{
  for (int64_t i = 0; i < nrows; i++)
    for (int64_t j = indptr[i]; j < indptr[i + 1]; j++) rows_out[j] = i;
}

/* ---- legate_sparse.random: host twin of the device generator (legate-sparse_b200/csrc/
 *      b2s_gallery.cu, random_fill_kernel).  No upstream counterpart (the reference's tests densify
 *      cupynumeric random arrays, tests/integration/utils/sample.py:21-45); this restatement is the
 *      checker of the device generator and the generator of the CPU arm's matrix in bench.py
 *      (parallel first touch: every thread initialises the rows it will later multiply). --------- */
static uint64_t ref_mix64(uint64_t t)
{
  t = (t ^ (t >> 30)) * 0xBF58476D1CE4E5B9ull;
  t = (t ^ (t >> 27)) * 0x94D049BB133111EBull;
  return t ^ (t >> 31);
}
static int64_t ref_extras_before(int64_t a, int64_t m, int64_t rem)
{
  return (a / m) * rem + ((a % m) < rem ? (a % m) : rem);
}
static int64_t ref_row_start(int64_t i, int64_t m, int64_t q, int64_t rem, int64_t shift)
{
  return i * q + ref_extras_before(i + shift, m, rem) - ref_extras_before(shift, m, rem);
}

#include <math.h>
/* rows [r0, r1) of the m x n matrix with nnz_total entries; indptr_local has r1-r0+1 entries
 * (0-based), crd/vals hold indptr_local[r1-r0] entries; values uniform in [lo, hi). */
void ref_random_csr_f64(int64_t m, int64_t n, int64_t nnz_total, uint64_t seed, int64_t r0, int64_t r1,
                        double lo, double hi, int64_t* indptr_local, int64_t* crd, double* vals)
{
  const int64_t q = nnz_total / m, rem = nnz_total % m;
  const int64_t shift = (int64_t)(ref_mix64(seed) % (uint64_t)m);
  const int64_t base0 = ref_row_start(r0, m, q, rem, shift);
#pragma omp parallel for schedule(static)
  for (int64_t i = r0; i < r1; i++) {
    const int64_t k    = q + (((i + shift) % m) < rem ? 1 : 0);
    const int64_t base = ref_row_start(i, m, q, rem, shift) - base0;
    indptr_local[i - r0] = base;
    if (i == r1 - 1) indptr_local[r1 - r0] = base + k;
    const uint64_t rowkey = ref_mix64(seed + 0x9E3779B97F4A7C15ull * (uint64_t)(i + 1));
    for (int64_t j = 0; j < k; j++) {
      const uint64_t s0 = (uint64_t)(((unsigned __int128)(uint64_t)j * (uint64_t)n) / (uint64_t)k);
      const uint64_t s1 = (uint64_t)(((unsigned __int128)(uint64_t)(j + 1) * (uint64_t)n) / (uint64_t)k);
      const uint64_t h  = ref_mix64(rowkey + (uint64_t)j);
      crd[base + j]     = (int64_t)(s0 + h % (s1 - s0));
      const uint64_t h2 = ref_mix64(h ^ 0x632BE59BD9B4E019ull);
      vals[base + j]    = fma(hi - lo, (double)(h2 >> 11) * (1.0 / 9007199254740992.0), lo);
    }
  }
  if (r1 == r0) indptr_local[0] = 0;
}

/* y and x for the CPU arm, first-touched in parallel like the matrix */
void ref_fill_uniform_f64(int64_t n, uint64_t seed, double* x)
{
#pragma omp parallel for schedule(static)
  for (int64_t i = 0; i < n; i++)
    x[i] = (double)(ref_mix64(seed + (uint64_t)i) >> 11) * (1.0 / 9007199254740992.0);
}
