// b2s_board.cuh — cross-GPU all-reduce(sum) of one scalar per rank through peer-mapped boards
// (see the comment block in b2s_vec.cu).  Shared by the stand-alone exchange kernel and by the
// kernels that fold the exchange into their final reduction (cg_update, the SpMV's fused dot).
#pragma once
#include "b2s_common.cuh"

namespace b2s {

constexpr int kBoardChannels = 4;
constexpr int kBoardRanks    = kMaxPeers + 1;
struct alignas(32) BoardSlot {
  unsigned char value[16];      // packed form (values <= 8 bytes): bytes 0..7 payload, 8..15 sequence number
  unsigned long long seq;       // c128 form: 16-byte value above, sequence number here
  unsigned long long pad;
};
struct BoardPtrs { BoardSlot* b[kBoardRanks]; };

__device__ __forceinline__ void st_release_sys(unsigned long long* p, unsigned long long v) {
  asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}

// 16-byte slot accesses: {payload, seq} travel in ONE store / load, so no fence is needed between
// the value and its flag (a 16-byte aligned vector access is a single transaction on NVLink and L2)
__device__ __forceinline__ void st_slot16(void* p, unsigned long long payload, unsigned long long seq) {
  asm volatile("st.volatile.global.v2.u64 [%0], {%1, %2};" ::"l"(p), "l"(payload), "l"(seq) : "memory");
}
__device__ __forceinline__ void ld_slot16(const void* p, unsigned long long* payload, unsigned long long* seq) {
  asm volatile("ld.volatile.global.v2.u64 {%0, %1}, [%2];" : "=l"(*payload), "=l"(*seq) : "l"(p) : "memory");
}

// what a kernel needs to finish its reduction with an exchange (nranks <= 1: no exchange)
template <typename V>
struct BoardArgs {
  BoardPtrs boards;
  int rank, nranks, channel;
  unsigned long long* seq_counters;
  V* cur_out;    // optional: prev_out[0] = cur_out[0]; cur_out[0] = sum
  V* prev_out;
  int* err;
};

// Executed by ONE full warp (32 converged lanes).  `mine`: this rank's partial (same in every lane).
// Returns the sum over the ranks in rank order (valid in lane 0) and performs the optional stores.
template <typename V>
__device__ __forceinline__ V board_exchange_warp(V mine, const BoardArgs<V>& bx, V* vals /* shared, >= kBoardRanks */) {
  const int t = threadIdx.x & 31;
  const unsigned long long seq = bx.seq_counters[bx.channel] + 1ull;
  const int slot_base = (bx.channel * 2 + (int)(seq & 1ull)) * kBoardRanks;
  if (t < bx.nranks) {
    BoardSlot* dst = bx.boards.b[t] + slot_base + bx.rank;        // my slot on rank t's board
    BoardSlot* src = bx.boards.b[bx.rank] + slot_base + t;        // rank t's slot on my board
    const long long t0 = clock64();
    bool ok = true;
    if constexpr (sizeof(V) <= 8) {
      unsigned long long payload = 0;
      memcpy(&payload, &mine, sizeof(V));
      st_slot16(dst, payload, seq);                               // slot bytes 0..15 = {payload, seq}
      unsigned long long got = 0, gseq = 0;
      while (true) {
        ld_slot16(src, &got, &gseq);
        if (gseq == seq) break;
        if (clock64() - t0 > (1ll << 34)) { ok = false; break; }  // ~8 s: a peer died — do not hang the GPU
      }
      V v;
      memcpy(&v, &got, sizeof(V));
      vals[t] = v;
    } else {
      *reinterpret_cast<V*>(dst->value) = mine;
      st_release_sys(&dst->seq, seq);                              // value first, then the flag
      while (ld_acquire_sys(&src->seq) != seq) {
        if (clock64() - t0 > (1ll << 34)) { ok = false; break; }
      }
      vals[t] = *reinterpret_cast<const V*>(src->value);           // ordered after the acquire load above
    }
    if (!ok && bx.err) atomicExch(bx.err, 1);
  }
  __syncwarp();
  V tot = zero_of<V>();
  if (t == 0) {
    tot = vals[0];
    for (int g = 1; g < bx.nranks; ++g) tot = vadd(tot, vals[g]);
    if (bx.prev_out) bx.prev_out[0] = bx.cur_out[0];
    if (bx.cur_out) bx.cur_out[0] = tot;
    bx.seq_counters[bx.channel] = seq;
  }
  return tot;
}

// host side: fill BoardArgs from the C ABI arguments (nranks <= 1 or boards == NULL: disabled)
template <typename V>
static inline int make_board_args(void* const* boards, int rank, int nranks, int channel, void* seq_counters,
                                  void* cur_out, void* prev_out, void* err, BoardArgs<V>* out) {
  BoardArgs<V> bx{};
  bx.nranks = 0;
  if (boards != nullptr && nranks > 1) {
    if (nranks > kBoardRanks || rank < 0 || rank >= nranks || channel < 0 || channel >= kBoardChannels || !seq_counters) {
      set_error("bad board arguments (rank %d of %d, channel %d)", rank, nranks, channel);
      return B2S_ERR_ARG;
    }
    for (int g = 0; g < nranks; ++g) {
      if (!boards[g]) { set_error("null board pointer"); return B2S_ERR_ARG; }
      bx.boards.b[g] = reinterpret_cast<BoardSlot*>(boards[g]);
    }
    bx.rank = rank; bx.nranks = nranks; bx.channel = channel;
    bx.seq_counters = reinterpret_cast<unsigned long long*>(seq_counters);
    bx.cur_out = reinterpret_cast<V*>(cur_out); bx.prev_out = reinterpret_cast<V*>(prev_out);
    bx.err = reinterpret_cast<int*>(err);
  }
  *out = bx;
  return B2S_OK;
}

}  // namespace b2s
