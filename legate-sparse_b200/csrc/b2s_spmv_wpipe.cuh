// b2s_spmv_wpipe.cuh — persistent TMA-fed CSR SpMV with WARP-AUTONOMOUS consumers.
//
// Included by b2s_spmv.cu.  Same producer protocol as spmv_pipe_kernel (one elected lane fills a
// shared-memory ring with TMA bulk copies of the tile's col / val / indptr slices); what changes
// is the consumer side, built for matrices whose x gathers go to L2 (no x window):
//
//   * the tile (1024 nnz) is cut into 8 SUB-TILES of 128 consecutive non-zeros, one per consumer
//     warp; the first row of every sub-tile comes from the plan (sub_row[], staged with the tile),
//     so no search is needed;
//   * a warp gathers x for its 128 pairs (4 independent gathers per lane), parks the products in
//     its own slice of the stage, __syncwarp()s, and reduces the rows of ITS sub-tile with 1..32
//     lanes per row — no CTA-wide barrier anywhere: a warp whose gathers are slow does not stall
//     the other seven;
//   * a row crossing sub-tiles: the sub-tile where the row STARTS owns y[r]; the others write
//     their piece to sub_head[] and spmv_subfixup_kernel adds the pieces in sub-tile order
//     (deterministic, no floating-point atomics).
#pragma once

namespace b2s {

constexpr int kSubPerTile = kPipeConsumers / 32;   // 8 sub-tiles (warps) per tile

template <typename V, typename I, int STAGES, bool DOT>
__global__ void __launch_bounds__(kPipeThreads)
spmv_wpipe_kernel(int64_t nrows, int64_t ncols, int64_t nnz, int64_t ntiles, int has_empty_rows,
                  const int64_t* __restrict__ indptr, const I* __restrict__ cols,
                  const V* __restrict__ vals, const V* __restrict__ x, V* __restrict__ y,
                  const int64_t* __restrict__ sub_row /* [ntiles*8 + 2] */,
                  V* __restrict__ sub_head, int64_t* __restrict__ sub_head_row,
                  V* __restrict__ dot_partials, const V* __restrict__ w) {
  constexpr int IPT = 4;
  using L = PipeLayout<V, I, IPT>;
  constexpr int T = L::T;       // 1024
  constexpr int W = 32 * IPT;   // 128
  // stage layout: vals | cols | rptr | meta | sub_row[10] (in the xwin slot)
  constexpr size_t STAGE = (L::xwin_off + 16 * 8 + 127) / 128 * 128;
  extern __shared__ __align__(128) unsigned char smem[];
  uint64_t* full_bar  = reinterpret_cast<uint64_t*>(smem + STAGE * STAGES);
  uint64_t* empty_bar = full_bar + STAGES;
  __shared__ V wsum[kSubPerTile];  // DOT only

  const int tid = threadIdx.x;
  if (tid == 0) {
    for (int s = 0; s < STAGES; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], kPipeConsumers); }
    fence_mbar_init();
  }
  __syncthreads();

  const uint64_t pol_stream = policy_evict_first();
  const uint64_t pol_keep   = policy_evict_last();

  if (tid >= kPipeConsumers) {
    // ============================== PRODUCER (one elected lane) ==============================
    if (tid == kPipeConsumers) {
      int64_t i = 0;
      for (int64_t t = blockIdx.x; t < ntiles; t += gridDim.x, ++i) {
        const int s = (int)(i % STAGES);
        const uint32_t ph = (uint32_t)((i / STAGES) & 1);
        mbar_wait(&empty_bar[s], ph ^ 1u);
        unsigned char* st = smem + STAGE * s;
        PipeMeta* meta = reinterpret_cast<PipeMeta*>(st + L::meta_off);
        const int64_t S = t * (int64_t)T;
        const int64_t E = min(S + (int64_t)T, nnz);
        const int64_t r_begin = sub_row[t * kSubPerTile];
        // row containing the last non-zero of the tile = first row of the next tile's first
        // sub-tile, unless that one starts exactly at a row boundary → conservative: use it
        const int64_t r_last = (t + 1 < ntiles) ? sub_row[(t + 1) * kSubPerTile] : nrows;
        const int64_t e  = min(r_last, nrows - 1) + 1;   // last indptr entry needed
        const int64_t ra = r_begin & ~(int64_t)1;
        int64_t n_ent = e - ra + 1;
        n_ent += (n_ent & 1);
        const bool rows_ok = (n_ent <= L::RCAP) && (ra + n_ent <= nrows + 1);
        const bool full = (E - S) == T;
        meta->r_begin = r_begin; meta->r_last = r_last; meta->ra = ra; meta->wbase = 0;
        meta->rows_staged = rows_ok; meta->win_staged = 0; meta->full_tile = full;
        uint32_t tx = 10 * 8;   // sub_row[t*8 .. t*8+9] (array is padded by 2 entries)
        if (full) tx += (uint32_t)(T * (sizeof(I) + sizeof(V)));
        if (rows_ok) tx += (uint32_t)(n_ent * 8);
        mbar_arrive_expect_tx(&full_bar[s], tx);
        tma_bulk_g2s(st + L::xwin_off, sub_row + t * kSubPerTile, 10 * 8, &full_bar[s], pol_stream);
        if (full) {
          tma_bulk_g2s(st + L::vals_off, vals + S, (uint32_t)(T * sizeof(V)), &full_bar[s], pol_stream);
          tma_bulk_g2s(st + L::cols_off, cols + S, (uint32_t)(T * sizeof(I)), &full_bar[s], pol_stream);
        }
        if (rows_ok) tma_bulk_g2s(st + L::rptr_off, indptr + ra, (uint32_t)(n_ent * 8), &full_bar[s], pol_stream);
      }
    }
    return;
  }

  // ================================== CONSUMERS (per warp) ==================================
  const int lane = tid & 31, warp = tid >> 5;
  int64_t i = 0;
  V dot_acc = zero_of<V>();
  for (int64_t t = blockIdx.x; t < ntiles; t += gridDim.x, ++i) {
    const int s = (int)(i % STAGES);
    const uint32_t ph = (uint32_t)((i / STAGES) & 1);
    mbar_wait(&full_bar[s], ph);
    unsigned char* st = smem + STAGE * s;
    V* svals = reinterpret_cast<V*>(st + L::vals_off);
    const I* scols = reinterpret_cast<const I*>(st + L::cols_off);
    const int64_t* srptr = reinterpret_cast<const int64_t*>(st + L::rptr_off);
    const int64_t* ssub = reinterpret_cast<const int64_t*>(st + L::xwin_off);
    const PipeMeta meta = *reinterpret_cast<const PipeMeta*>(st + L::meta_off);
    const int64_t S = t * (int64_t)T;
    const int64_t E = min(S + (int64_t)T, nnz);
    const int64_t r_max = min(meta.r_last, nrows - 1);
    const int64_t* rp = meta.rows_staged ? (srptr - meta.ra) : indptr;

    // empty rows strictly inside the tile: y = 0 (no barrier needed: disjoint from the row sums)
    if (has_empty_rows) {
      const int64_t r_hi = min(meta.r_last, nrows);
      for (int64_t r = meta.r_begin + tid; r < r_hi; r += kPipeConsumers)
        if (rp[r + 1] == rp[r]) y[r] = zero_of<V>();
    }

    const int64_t P0 = S + (int64_t)warp * W;
    const int64_t P1 = min(P0 + (int64_t)W, E);
    const int64_t sub = t * kSubPerTile + warp;
    if (P0 >= E) {                    // sub-tile beyond the end of the (last, partial) tile
      if (lane == 0) sub_head_row[sub] = -1;
      mbar_arrive(&empty_bar[s]);
      continue;
    }
    // ---- products of this warp's 128 pairs → its slice of the stage ----
    V* wprod = svals + warp * W;
    if (meta.full_tile) {
      const int q = warp * W + lane * IPT;
      I c[IPT];
      V a[IPT];
      memcpy(c, scols + q, sizeof(I) * IPT);   // LDS.128
      memcpy(a, svals + q, sizeof(V) * IPT);
      V xv[IPT];
#pragma unroll
      for (int k = 0; k < IPT; ++k) xv[k] = ld_gather<V>(x + (int64_t)c[k], pol_keep);
#pragma unroll
      for (int k = 0; k < IPT; ++k) a[k] = vmul(a[k], xv[k]);
      memcpy(svals + q, a, sizeof(V) * IPT);   // STS.128
    } else {
      for (int64_t p = P0 + lane; p < P1; p += 32) {
        const int64_t c = (int64_t)ld_stream<I>(cols + p, pol_stream);
        const V a = ld_stream<V>(vals + p, pol_stream);
        wprod[p - P0] = vmul(a, ld_gather<V>(x + c, pol_keep));
      }
    }
    __syncwarp();

    // ---- rows of the sub-tile: [r_lo, r_hi] (non-empty rows containing P0 and P1-1) ----
    const int64_t r_lo = ssub[warp];
    int64_t r_hi = ssub[warp + 1];             // row containing P1 (or the sentinel)
    if (r_hi > r_max) r_hi = r_max;
    while (r_hi > r_lo && rp[r_hi] >= P1) --r_hi;   // step back to the row containing P1-1
    const int64_t nr = r_hi - r_lo + 1;
    const bool has_head = rp[r_lo] < P0;
    if (!has_head && lane == 0) sub_head_row[sub] = -1;
    int lanes = 1;
    while (lanes < 32 && (int64_t)(lanes * 2) * nr <= 32) lanes <<= 1;
    const int groups = 32 / lanes;
    const int gl = lane & (lanes - 1);
    for (int64_t base = 0; base < nr; base += groups) {
      const int64_t r = r_lo + base + lane / lanes;
      const bool valid = r <= r_hi;
      int64_t lo_g = 0, hi_g = 0;
      if (valid) { lo_g = rp[r]; hi_g = rp[r + 1]; }
      const int64_t lo = max(lo_g, P0), hi = min(hi_g, P1);
      V sum = zero_of<V>();
      for (int64_t p = lo + gl; p < hi; p += lanes) sum = vadd(sum, wprod[p - P0]);
      sum = group_reduce(sum, lanes);
      if (valid && gl == 0 && hi_g > lo_g) {
        if (lo_g < P0) {
          sub_head[sub] = sum;                                  // continues a row of an earlier sub-tile
          sub_head_row[sub] = r * 2 + (lo_g >= P0 - W ? 1 : 0);
        } else {
          y[r] = sum;                                           // this sub-tile owns y[r]
        }
        if (DOT) dot_acc = vfma(w[r], sum, dot_acc);
      }
    }
    mbar_arrive(&empty_bar[s]);  // this thread is done with the stage
  }
  if (DOT) {
    V sacc = dot_acc;
    for (int o = 16; o > 0; o >>= 1) sacc = vadd(sacc, vshfl_xor(sacc, o));
    if (lane == 0) wsum[warp] = sacc;
    consumer_bar_sync();
    if (tid == 0) {
      V tot = wsum[0];
      for (int k = 1; k < kSubPerTile; ++k) tot = vadd(tot, wsum[k]);
      dot_partials[blockIdx.x] = tot;
    }
  }
}

}  // namespace b2s
