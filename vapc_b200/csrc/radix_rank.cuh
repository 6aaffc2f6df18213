// Tile-level machinery shared by the radix-sort passes (sort.cu) and the record bucketing pass (bucket.cu):
// stable ranking of 8-bit digits inside a 256-thread CTA, publication of the tile's digit counts and the
// decoupled look-back that turns them into global offsets.
//
// Ranking: the lanes of a warp holding the same digit find each other through a shared-memory word per
// (warp, digit) that they OR their lane bit into (one ATOMS + one LDS per key).  The two alternatives were
// measured on B200 (profiles/): match.any runs on the SM's single ADU pipe at a cost proportional to the
// distinct values in the warp (sort kernels 99 % ADU-bound), and peer masks from 8 vote.ballot's cost ~35 ALU
// instructions per key (0.65 ms per pass at 1e8 keys, ALU-bound).  With the shared-memory masks the kernels
// are bound by shared-memory wavefronts, so every shared access below is there once per key at most.
#pragma once
#include "common.cuh"

namespace rr {
constexpr int kRadixBits = 8;
constexpr int kRadix = 1 << kRadixBits;
constexpr int kThreads = 256;  // == kRadix: thread t owns digit t in the per-digit steps
constexpr int kWarps = kThreads / 32;
#ifndef VAPC_SORT_LOOK
#define VAPC_SORT_LOOK 4
#endif
constexpr int kLookWindow = VAPC_SORT_LOOK;  // look-back words fetched at once (a hop is an L2 round trip)

// look[tile][digit] packs a 2-bit status with a count:
//   0 = not ready, 1 = this tile's own count, 2 = inclusive count of all tiles of the chain up to this one.
// 32-bit words hold 30-bit counts (n < 2^30); 64-bit words are used beyond that.
template <typename LookT> struct Look;
template <> struct Look<uint32_t> {
  static constexpr int kShift = 30;
  static __device__ __forceinline__ uint32_t load(const uint32_t* p) {
    uint32_t v;
    asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
  }
  static __device__ __forceinline__ void store(uint32_t* p, uint32_t v) {
    asm volatile("st.relaxed.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
  }
};
template <> struct Look<unsigned long long> {
  static constexpr int kShift = 62;
  static __device__ __forceinline__ unsigned long long load(const unsigned long long* p) {
    unsigned long long v;
    asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
  }
  static __device__ __forceinline__ void store(unsigned long long* p, unsigned long long v) {
    asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
  }
};

typedef unsigned MaskBuffers[3][kWarps][kRadix];  // ranking scratch, triple-buffered (24 KB)

struct RankSmem {
  unsigned warp_cnt[kWarps][kRadix];  // per-warp running digit counters -> tile position of each warp's first key of a digit
  unsigned bin_start[kRadix];         // tile-local first slot of each digit
  unsigned delta[kRadix];             // global slot of the key at tile position p with digit d = p + delta[d]
  unsigned warp_s[33];
  unsigned tile;
};

// Zeroes the running counters and the mask words with 128-bit stores: 8 instructions per thread instead of 32 (the word-wise version
// showed up as MIO-queue throttle — 5.6 % of the scatter kernel's stall samples, as many store instructions as its whole ranking).
// Both arrays start 16-byte aligned (warp_cnt leads RankSmem, the masks lead the aligned union of the kernels' shared structs).
__device__ __forceinline__ void rank_init(RankSmem& sm, MaskBuffers& masks) {
  const int tid = threadIdx.x;
  uint4* wc4 = reinterpret_cast<uint4*>(&sm.warp_cnt[0][0]);
  uint4* mk4 = reinterpret_cast<uint4*>(&masks[0][0][0]);
  const uint4 zero = make_uint4(0u, 0u, 0u, 0u);
#pragma unroll
  for (int i = 0; i < kWarps * kRadix / 4 / kThreads; ++i) wc4[i * kThreads + tid] = zero;
#pragma unroll
  for (int i = 0; i < 3 * kWarps * kRadix / 4 / kThreads; ++i) mk4[i * kThreads + tid] = zero;
}

// Stable rank of ITEMS keys per lane inside the warp; digit(k) is the digit of the lane's k-th key and step k
// of all lanes precedes step k+1 in input order.  rank2 receives two 16-bit ranks per word: the number of keys
// of the same digit that this warp holds in front of the key.  The lowest lane of each peer set advances the
// warp's running digit counter with one shared-memory atomic and clears the mask word one step later (hence
// three mask buffers and a single warp barrier per step).
template <int ITEMS, typename DigitFn>
__device__ __forceinline__ void warp_rank(DigitFn digit, RankSmem& sm, MaskBuffers& masks, unsigned (&rank2)[ITEMS / 2]) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const unsigned lt = (1u << lane) - 1u;
  unsigned* wc = sm.warp_cnt[warp];
  unsigned prev_d = 0;
  bool prev_leader = false;
#pragma unroll
  for (int k = 0; k < ITEMS; ++k) {
    unsigned* mk = masks[k % 3][warp];
    const unsigned d = digit(k);
    atomicOr(&mk[d], 1u << lane);
    __syncwarp();
    if (k > 0 && prev_leader) masks[(k + 2) % 3][warp][prev_d] = 0;
    const unsigned peers = mk[d];
    const int leader = __ffs(peers) - 1;
    const bool is_leader = lane == leader;
    unsigned before = 0;
    if (is_leader) before = atomicAdd(&wc[d], (unsigned)__popc(peers));
    before = __shfl_sync(0xffffffffu, before, leader);
    const unsigned r = before + (unsigned)__popc(peers & lt);
    rank2[k / 2] = (k & 1) ? (rank2[k / 2] | (r << 16)) : r;
    prev_d = d;
    prev_leader = is_leader;
  }
}
__device__ __forceinline__ unsigned rank_of(const unsigned* rank2, int k) { return (k & 1) ? (rank2[k / 2] >> 16) : (rank2[k / 2] & 0xffffu); }

// Call after a __syncthreads() that follows warp_rank.  Thread tid == digit: publishes the tile's count of its
// digit for the tiles behind (inclusive at once when the tile is the first of its chain), computes the digit's
// first slot inside the tile and turns warp_cnt[w][digit] into each warp's first slot.  Returns the count;
// *bstart = first slot.  Ends with a barrier.
template <typename LookT>
__device__ __forceinline__ unsigned digit_offsets(RankSmem& sm, LookT* mine, bool first_in_chain, unsigned* bstart) {
  const int tid = threadIdx.x;
  unsigned cnt[kWarps];
  unsigned run = 0;
#pragma unroll
  for (int w = 0; w < kWarps; ++w) {
    cnt[w] = sm.warp_cnt[w][tid];
    run += cnt[w];
  }
  Look<LookT>::store(mine, (LookT)run | ((LookT)(first_in_chain ? 2 : 1) << Look<LookT>::kShift));
  unsigned total;
  const unsigned b = block_exclusive_scan(run, sm.warp_s, &total);
  sm.bin_start[tid] = b;
  unsigned acc = b;
#pragma unroll
  for (int w = 0; w < kWarps; ++w) {
    sm.warp_cnt[w][tid] = acc;
    acc += cnt[w];
  }
  __syncthreads();
  *bstart = b;
  return run;
}

// Decoupled look-back (thread tid == digit): the number of keys of this digit in the `in_front` tiles that
// precede this one in its chain (their words lie directly in front of `mine`, kRadix words apart).  Sums the
// tiles' own counts until one has published its inclusive count, then publishes this tile's inclusive count.
// kLookWindow words are fetched at once; words behind the first inclusive one are simply ignored, a word that
// is not ready yet is polled again.  Tiles take their place in the chain from an atomic ticket, so every tile
// in front is resident or done and the poll cannot starve.
template <typename LookT>
__device__ __forceinline__ LookT look_back(LookT* mine, int64_t in_front, unsigned run) {
  LookT excl = 0;
  if (in_front > 0) {
    const LookT vmask = ((LookT)1 << Look<LookT>::kShift) - 1;
    int64_t back = 1;  // distance of the first word of the window
    bool done = false;
    while (!done) {
      LookT w[kLookWindow];
#pragma unroll
      for (int j = 0; j < kLookWindow; ++j)
        w[j] = (back + j <= in_front) ? Look<LookT>::load(mine - (back + j) * kRadix) : ((LookT)2 << Look<LookT>::kShift);
#pragma unroll
      for (int j = 0; j < kLookWindow; ++j) {
        if (done) break;
        LookT v = w[j];
        while ((v >> Look<LookT>::kShift) == 0) {
#ifdef VAPC_POLL_SLEEP
          __nanosleep(VAPC_POLL_SLEEP);
#endif
          v = Look<LookT>::load(mine - (back + j) * kRadix);
        }
        excl += v & vmask;
        done = (v >> Look<LookT>::kShift) == 2;
      }
      back += kLookWindow;
    }
    Look<LookT>::store(mine, (excl + run) | ((LookT)2 << Look<LookT>::kShift));
  }
  return excl;
}

// hist[blockIdx.x][.] -> exclusive scan over the digits (+ base[blockIdx.x % nbase] if given): the first output
// slot of each digit.  One CTA per histogram.
static __global__ void __launch_bounds__(kThreads) digit_scan_kernel(uint32_t* __restrict__ hist, const uint32_t* __restrict__ base, int nbase) {
  __shared__ unsigned warp_s[33];
  unsigned total;
  uint32_t* h = hist + (int64_t)blockIdx.x * kRadix;
  const unsigned v = h[threadIdx.x];
  h[threadIdx.x] = block_exclusive_scan(v, warp_s, &total) + (base ? base[blockIdx.x % nbase] : 0u);
}
}  // namespace rr
