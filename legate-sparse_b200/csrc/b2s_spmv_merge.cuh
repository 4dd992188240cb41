// b2s_spmv_merge.cuh — persistent, TMA-fed, warp-autonomous CSR SpMV (the default kernel).
//
// Included by b2s_spmv.cu after b2s_spmv_pipe.cuh (shares PipeLayout / PipeMeta / the producer
// protocol).  Same smem ring filled by one producer lane with TMA bulk copies (UBLKCP) of the
// tile's col / val / indptr slices (+ the x window for banded matrices); the consumer side is
// nnz-balanced down to the LANE:
//
//   * every consumer warp owns a SUB-TILE of 32*IPT consecutive non-zeros of the staged tile;
//     lane l holds IPT consecutive (col,val) pairs (one LDS.128 for 4 int32 cols), issues its
//     IPT x-gathers back to back (all independent → maximal memory-level parallelism), and
//     accumulates its products while walking the row boundaries it finds by one binary search
//     of the staged indptr slice;
//   * rows that start and end inside a lane are written at once; the open pieces are combined
//     by ONE warp-level segmented scan (5 shuffle steps, fixed order → deterministic);
//   * a row that crosses sub-tiles: the sub-tile where it STARTS owns y[r]; later sub-tiles
//     write their piece to head[sub] (+ the row id), and spmv_subfixup_kernel adds the pieces
//     in sub-tile order.  No floating-point atomics, no CTA-wide barrier, perfect balance for
//     any row-length distribution (power-law rows included).
//   * empty rows are zero-filled by a strided pre-pass over the staged indptr slice (skipped
//     when the plan found no empty row).
#pragma once

namespace b2s {

// (value, flag) segmented-scan combine: flag marks "a row ended at/after this lane's start"
template <typename V>
struct SegPair { V v; int f; };

template <typename V>
__device__ __forceinline__ SegPair<V> seg_shfl_up(SegPair<V> a, int d) {
  SegPair<V> r;
  r.v = vshfl_up(a.v, d);
  r.f = __shfl_up_sync(0xffffffffu, a.f, d);
  return r;
}

template <typename V, typename I, int IPT, int STAGES, bool WINDOW, bool DOT>
__global__ void __launch_bounds__(kPipeThreads)
spmv_merge_kernel(int64_t nrows, int64_t ncols, int64_t nnz, int64_t ntiles, int has_empty_rows,
                  const int64_t* __restrict__ indptr, const I* __restrict__ cols,
                  const V* __restrict__ vals, const V* __restrict__ x, V* __restrict__ y,
                  const int64_t* __restrict__ tile_row, const int64_t* __restrict__ tile_win,
                  V* __restrict__ sub_head, int64_t* __restrict__ sub_head_row,
                  V* __restrict__ dot_partials, const V* __restrict__ w) {
  using L = PipeLayout<V, I, IPT>;
  constexpr int T = L::T;                 // 256 * IPT
  constexpr int W = 32 * IPT;             // non-zeros per warp sub-tile
  constexpr int CW = kPipeConsumers / 32; // consumer warps
  constexpr size_t STAGE = L::stage_bytes(WINDOW);
  extern __shared__ __align__(128) unsigned char smem[];
  uint64_t* full_bar  = reinterpret_cast<uint64_t*>(smem + STAGE * STAGES);
  uint64_t* empty_bar = full_bar + STAGES;
  __shared__ V wsum[CW];  // DOT only

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
        const int64_t r_begin = tile_row[t], r_last = tile_row[t + 1];
        uint32_t tx = 0;
        const int64_t e  = min(r_last, nrows - 1) + 1;   // last indptr entry needed
        const int64_t ra = r_begin & ~(int64_t)1;
        int64_t n_ent = e - ra + 1;
        n_ent += (n_ent & 1);
        const bool rows_ok = (n_ent <= L::RCAP) && (ra + n_ent <= nrows + 1);
        const bool full = (E - S) == T;
        int64_t wbase = 0, wcnt = 0;
        bool win_ok = false;
        uint32_t win_bytes = 0;
        if (WINDOW) {
          wbase = tile_win[2 * t];
          wcnt  = tile_win[2 * t + 1];
          if (wcnt > 0 && wcnt <= kPipeWinCap) {
            constexpr int PER16 = (16 / (int)sizeof(V)) > 0 ? (16 / (int)sizeof(V)) : 1;
            int64_t want = (wcnt + PER16 - 1) / PER16 * PER16;
            if (wbase + want <= ncols) { win_ok = true; win_bytes = (uint32_t)(want * sizeof(V)); }
          }
        }
        meta->r_begin = r_begin; meta->r_last = r_last; meta->ra = ra; meta->wbase = wbase;
        meta->rows_staged = rows_ok; meta->win_staged = win_ok; meta->full_tile = full;
        if (full) tx += (uint32_t)(T * (sizeof(I) + sizeof(V)));
        if (rows_ok) tx += (uint32_t)(n_ent * 8);
        tx += win_bytes;
        if (tx > 0) {
          mbar_arrive_expect_tx(&full_bar[s], tx);
          if (full) {
            tma_bulk_g2s(st + L::vals_off, vals + S, (uint32_t)(T * sizeof(V)), &full_bar[s], pol_stream);
            tma_bulk_g2s(st + L::cols_off, cols + S, (uint32_t)(T * sizeof(I)), &full_bar[s], pol_stream);
          }
          if (rows_ok) tma_bulk_g2s(st + L::rptr_off, indptr + ra, (uint32_t)(n_ent * 8), &full_bar[s], pol_stream);
          if (win_ok) tma_bulk_g2s(st + L::xwin_off, x + wbase, win_bytes, &full_bar[s], pol_keep);
        } else {
          mbar_arrive(&full_bar[s]);
        }
      }
    }
    return;
  }

  // ================================== CONSUMERS ==================================
  const int lane = tid & 31, warp = tid >> 5;
  int64_t i = 0;
  V dot_acc = zero_of<V>();
  for (int64_t t = blockIdx.x; t < ntiles; t += gridDim.x, ++i) {
    const int s = (int)(i % STAGES);
    const uint32_t ph = (uint32_t)((i / STAGES) & 1);
    mbar_wait(&full_bar[s], ph);
    unsigned char* st = smem + STAGE * s;
    const V* svals = reinterpret_cast<const V*>(st + L::vals_off);
    const I* scols = reinterpret_cast<const I*>(st + L::cols_off);
    const int64_t* srptr = reinterpret_cast<const int64_t*>(st + L::rptr_off);
    const V* sxwin = reinterpret_cast<const V*>(st + L::xwin_off);
    const PipeMeta meta = *reinterpret_cast<const PipeMeta*>(st + L::meta_off);
    const int64_t S = t * (int64_t)T;
    const int64_t E = min(S + (int64_t)T, nnz);
    const bool use_win = WINDOW && meta.win_staged;
    const int64_t r_begin = meta.r_begin;
    const int64_t r_max = min(meta.r_last, nrows - 1);   // last real row touched by the tile
    // row pointer accessor: shared-memory slice when staged, global otherwise
    const int64_t* rp = meta.rows_staged ? (srptr - meta.ra) : indptr;

    // ---- empty rows strictly inside the tile: y = 0 ----
    if (has_empty_rows) {
      const int64_t r_hi = min(meta.r_last, nrows);  // rows [r_begin, r_hi) are "inside"
      for (int64_t r = r_begin + tid; r < r_hi; r += kPipeConsumers)
        if (rp[r + 1] == rp[r]) y[r] = zero_of<V>();
    }

    // ---- this lane's IPT consecutive non-zeros ----
    const int64_t P0 = S + (int64_t)warp * W;       // first nnz of the warp's sub-tile
    const int64_t e0 = P0 + (int64_t)lane * IPT;    // first nnz of this lane
    const int64_t sub = t * CW + warp;              // global sub-tile id
    V prod[IPT];
    if (e0 < E) {
      I c[IPT];
      V a[IPT];
      if (meta.full_tile) {
        const int q = (int)(e0 - S);
        memcpy(c, scols + q, sizeof(I) * IPT);   // LDS.128
        memcpy(a, svals + q, sizeof(V) * IPT);
      } else {
#pragma unroll
        for (int k = 0; k < IPT; ++k) {
          const int64_t p = e0 + k < E ? e0 + k : E - 1;   // clamp (masked below)
          c[k] = ld_stream<I>(cols + p, pol_stream);
          a[k] = ld_stream<V>(vals + p, pol_stream);
        }
      }
      V xv[IPT];
#pragma unroll
      for (int k = 0; k < IPT; ++k) {
        if (use_win) xv[k] = sxwin[(int64_t)c[k] - meta.wbase];
        else         xv[k] = ld_gather<V>(x + (int64_t)c[k], pol_keep);
      }
#pragma unroll
      for (int k = 0; k < IPT; ++k) prod[k] = (e0 + k < E) ? vmul(a[k], xv[k]) : zero_of<V>();
    } else {
#pragma unroll
      for (int k = 0; k < IPT; ++k) prod[k] = zero_of<V>();
    }

    // ---- row of the lane's first element: last r in [r_begin, r_max] with rp[r] <= e0 ----
    int64_t row = r_begin;
    int64_t row_end = 0;
    const bool active = e0 < E;
    if (active) {
      int64_t lo = r_begin, hi = r_max + 1;       // first index in (lo,hi] with rp[idx] > e0
      while (lo < hi) {
        const int64_t mid = (lo + hi) >> 1;
        if (rp[mid] <= e0) lo = mid + 1; else hi = mid;
      }
      row = lo - 1;
      row_end = rp[row + 1];
    }

    // ---- walk: pieces closed inside the lane are written; keep first closed piece + open tail ----
    V first_sum = zero_of<V>();   // sum up to the first row end inside the lane
    int64_t first_row = -1;
    V acc = zero_of<V>();
    int has_end = 0;
    if (active) {
#pragma unroll
      for (int k = 0; k < IPT; ++k) {
        const int64_t e = e0 + k;
        if (e < E) {
          acc = vadd(acc, prod[k]);
          if (e + 1 == row_end) {          // e is the last non-zero of `row`
            if (!has_end) { first_sum = acc; first_row = row; has_end = 1; }
            else {
              // the row started inside this lane → complete → this sub-tile owns it
              y[row] = acc;
              if (DOT) dot_acc = vfma(w[row], acc, dot_acc);
            }
            acc = zero_of<V>();
            // next non-empty row
            ++row;
            while (row <= r_max && rp[row + 1] == rp[row]) ++row;
            row_end = row <= r_max ? rp[row + 1] : INT64_MAX;
          }
        }
      }
    }
    // lane summary: v = open tail (if a row ended inside) else the whole lane sum
    SegPair<V> me;
    me.v = acc;
    me.f = has_end;
    // inclusive segmented scan over lanes:  (a,fa) ⊕ (b,fb) = (fb ? b : a+b, fa|fb)
    SegPair<V> inc = me;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      SegPair<V> up = seg_shfl_up(inc, d);
      if (lane >= d) {
        if (!inc.f) inc.v = vadd(up.v, inc.v);
        inc.f |= up.f;
      }
    }
    // exclusive = previous lane's inclusive
    SegPair<V> exc = seg_shfl_up(inc, 1);
    if (lane == 0) { exc.v = zero_of<V>(); exc.f = 0; }

    // the row of the warp's first element (lane 0) started before the sub-tile?
    const int64_t row0 = __shfl_sync(0xffffffffu, row, 0);  // valid when lane 0 is active
    bool wrote_head = false;
    if (active && has_end) {
      const V total = vadd(exc.v, first_sum);
      const int64_t rstart = rp[first_row];
      if (!exc.f && rstart < P0) {
        // first closed piece of the warp and the row began in an earlier sub-tile → head piece
        sub_head[sub] = total;
        sub_head_row[sub] = first_row * 2 + (rstart >= P0 - W ? 1 : 0);
        wrote_head = true;
        if (DOT) dot_acc = vfma(w[first_row], total, dot_acc);
      } else {
        y[first_row] = total;   // row starts inside this sub-tile (owner) → complete
        if (DOT) dot_acc = vfma(w[first_row], total, dot_acc);
      }
    }
    // open tail at the end of the sub-tile (lane 31's inclusive value) — or the whole sub-tile
    const int any_end = __shfl_sync(0xffffffffu, inc.f, 31);
    const V tail_val = vshfl_idx(inc.v, 31);
    const int64_t tail_row = __shfl_sync(0xffffffffu, row, 31);
    const int64_t sub_end = min(P0 + (int64_t)W, E);
    const unsigned head_mask = __ballot_sync(0xffffffffu, wrote_head);
    if (lane == 31) {
      bool head_written = head_mask != 0;
      if (P0 < E) {
        // is there an open piece?  (the last element of the sub-tile does not close its row)
        const bool open = (tail_row <= r_max) && (rp[tail_row] < sub_end) && (rp[tail_row + 1] > sub_end);
        if (open) {
          const int64_t rstart = rp[tail_row];
          if (!any_end && rstart < P0) {
            // the row covers the whole sub-tile and started earlier → the whole sum is a head piece
            sub_head[sub] = tail_val;
            sub_head_row[sub] = tail_row * 2 + (rstart >= P0 - W ? 1 : 0);
            head_written = true;
            if (DOT) dot_acc = vfma(w[tail_row], tail_val, dot_acc);
          } else {
            y[tail_row] = tail_val;   // owner writes its piece; later sub-tiles add heads
            if (DOT) dot_acc = vfma(w[tail_row], tail_val, dot_acc);
          }
        }
      }
      if (!head_written) sub_head_row[sub] = -1;
    }
    (void)row0;
    mbar_arrive(&empty_bar[s]);  // this thread is done with the stage
  }
  if (DOT) {
    V sacc = dot_acc;
    for (int o = 16; o > 0; o >>= 1) sacc = vadd(sacc, vshfl_xor(sacc, o));
    if (lane == 0) wsum[warp] = sacc;
    consumer_bar_sync();
    if (tid == 0) {
      V tot = wsum[0];
      for (int k = 1; k < CW; ++k) tot = vadd(tot, wsum[k]);
      dot_partials[blockIdx.x] = tot;
    }
  }
}

// y[r] += head pieces of the sub-tiles that continue row r, in sub-tile order (deterministic).
// One thread per sub-tile; only the FIRST continuation of a row (flag bit 0) does the work.
template <typename V>
__global__ void spmv_subfixup_kernel(int64_t nsub, const int64_t* __restrict__ sub_head_row,
                                     const V* __restrict__ sub_head, V* __restrict__ y) {
  int64_t u = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (u < 1 || u >= nsub) return;
  const int64_t code = sub_head_row[u];
  if (code < 0 || !(code & 1)) return;
  const int64_t r = code >> 1;
  V acc = y[r];
  for (int64_t v = u; v < nsub; ++v) {
    const int64_t cv = sub_head_row[v];
    if (cv < 0 || (cv >> 1) != r) break;
    acc = vadd(acc, sub_head[v]);
  }
  y[r] = acc;
}

}  // namespace b2s
