// In-house stable LSD radix sort of (key, point index) pairs — 8-bit digits, ONE sweep per pass.
//
//   digit_hist_kernel   one read of the keys -> the global digit histogram of EVERY pass (shared-memory
//                       atomics), digit_scan_kernel turns each into first-slot offsets          K B/elt
//   onesweep_kernel     per pass; tile = 4096 consecutive elements, one CTA of 256 threads:
//                       rank inside the tile, publish the tile's digit counts, decoupled look-back over
//                       the tiles in front for the global offsets, reorder the tile in shared memory and
//                       write each digit's run coalesced        reads K+4, writes K+4 B/elt, nothing else
// Ranking: the lanes of a warp holding the same digit find each other through a shared-memory word per
// (warp, digit) that they OR their lane bit into (one ATOMS + one LDS per key).  The two alternatives were
// measured on B200 (profiles/): match.any runs on the SM's single ADU pipe at a cost proportional to the
// distinct values in the warp (both sort kernels 99 % ADU-bound), and peer masks from 8 vote.ballot's cost
// ~35 ALU instructions per key (sort kernels ALU-bound, 0.65 ms per pass at 1e8 keys).
// Ranking keeps equal digits in input order, so every pass — and therefore the whole sort — is
// stable: points of one voxel stay in original index order (tie-break "lowest original index"
// and "first point of a voxel" fall out of that; see include/vapc_b200.h).
#include <algorithm>

#include "common.cuh"

namespace {
constexpr int kRadixBits = 8;
constexpr int kRadix = 1 << kRadixBits;
constexpr int kThreads = 256;  // == kRadix: thread t owns digit t in the per-digit steps
constexpr int kWarps = kThreads / 32;
constexpr int kItems = 16;
constexpr int kTile = kThreads * kItems;  // 4096
constexpr int kMaxPasses = 8;             // 64-bit keys
#ifndef VAPC_SORT_LOOK
#define VAPC_SORT_LOOK 4
#endif
constexpr int kLookWindow = VAPC_SORT_LOOK;  // look-back words fetched at once

template <typename KeyT>
__device__ __forceinline__ unsigned digit_of(KeyT k, int shift, unsigned mask) {
  return (unsigned)(k >> shift) & mask;
}

// ---- all digit histograms in one sweep over the keys ---------------------------------------------
// ghist[p][d] = number of keys whose p-th 8-bit digit is d.  Shared-memory atomics (no return value
// needed for counting), one flush of non-zero counters per CTA.
template <typename KeyT>
__global__ void __launch_bounds__(kThreads) digit_hist_kernel(const KeyT* __restrict__ keys, int64_t n, int passes, int end_bit,
                                                              uint32_t* __restrict__ ghist) {
  __shared__ unsigned sh[kMaxPasses][kRadix];
  for (int p = 0; p < passes; ++p) sh[p][threadIdx.x] = 0;
  __syncthreads();
  constexpr int kVec = 16 / sizeof(KeyT);  // keys per 16-byte load
  const int64_t nvec = n / kVec;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  const unsigned last_mask = (1u << (end_bit - (passes - 1) * kRadixBits)) - 1u;
  for (int64_t v = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; v < nvec; v += stride) {
    KeyT k[kVec];
    if (sizeof(KeyT) == 4) {
      const uint4 q = __ldcs(reinterpret_cast<const uint4*>(keys) + v);
      k[0] = (KeyT)q.x; k[1] = (KeyT)q.y; k[2] = (KeyT)q.z; k[3] = (KeyT)q.w;
    } else {
      const ulonglong2 q = __ldcs(reinterpret_cast<const ulonglong2*>(keys) + v);
      k[0] = (KeyT)q.x; k[1] = (KeyT)q.y;
    }
#pragma unroll
    for (int j = 0; j < kVec; ++j) {
      for (int p = 0; p < passes; ++p) {
        const unsigned m = p == passes - 1 ? last_mask : (unsigned)(kRadix - 1);
        atomicAdd(&sh[p][digit_of(k[j], p * kRadixBits, m)], 1u);
      }
    }
  }
  if (blockIdx.x == 0) {  // tail (n not a multiple of the vector width)
    for (int64_t i = nvec * kVec + threadIdx.x; i < n; i += blockDim.x) {
      const KeyT k = keys[i];
      for (int p = 0; p < passes; ++p) {
        const unsigned m = p == passes - 1 ? last_mask : (unsigned)(kRadix - 1);
        atomicAdd(&sh[p][digit_of(k, p * kRadixBits, m)], 1u);
      }
    }
  }
  __syncthreads();
  for (int p = 0; p < passes; ++p) {
    const unsigned c = sh[p][threadIdx.x];
    if (c) atomicAdd(&ghist[p * kRadix + threadIdx.x], c);
  }
}

// ghist[p][.] -> exclusive scan over the digits (first output slot of each digit); one CTA per pass
__global__ void __launch_bounds__(kThreads) digit_scan_kernel(uint32_t* __restrict__ ghist) {
  __shared__ unsigned warp_s[33];
  unsigned total;
  uint32_t* h = ghist + (int64_t)blockIdx.x * kRadix;
  const unsigned v = h[threadIdx.x];
  h[threadIdx.x] = block_exclusive_scan(v, warp_s, &total);
}

// ---- one radix pass in one sweep: rank + decoupled look-back + scatter --------------------------------
// Tiles are handed out by an atomic ticket, so a tile's predecessors are always resident or done and the
// look-back spin cannot starve.  look[tile][digit] packs a 2-bit status with the count:
//   0 = not ready, 1 = this tile's own count, 2 = inclusive count of all tiles up to this one.
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

template <typename KeyT> struct PairOf;  // (key, payload) as stored in the reorder buffer
template <> struct PairOf<uint32_t> { typedef uint2 type; };           // one 64-bit shared-memory access per element
template <> struct PairOf<unsigned long long> { typedef ulonglong2 type; };  // key + zero-extended payload

template <typename KeyT>
struct SweepSmem {
  unsigned warp_cnt[kWarps][kRadix];  // per-warp running digit counters -> tile position of each warp's first key of a digit
  unsigned bin_start[kRadix];         // tile-local first slot of each digit
  unsigned delta[kRadix];             // global slot of tile position p holding digit d = p + delta[d]
  unsigned warp_s[33];
  unsigned tile;
  alignas(16) union {
    unsigned masks[3][kWarps][kRadix];  // ranking: lanes of a warp holding each digit (triple-buffered)
    typename PairOf<KeyT>::type out[kTile];  // the tile reordered by digit
  } u;
};

template <typename KeyT, typename LookT, bool IOTA>
__global__ void __launch_bounds__(kThreads, sizeof(KeyT) == 4 ? 4 : 3) onesweep_kernel(const KeyT* __restrict__ keys_in, const uint32_t* __restrict__ idx_in,
                                                               KeyT* __restrict__ keys_out, uint32_t* __restrict__ idx_out, int64_t n,
                                                               int shift, unsigned mask, const uint32_t* __restrict__ digit_base,
                                                               LookT* __restrict__ look, unsigned* __restrict__ ticket) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  SweepSmem<KeyT>& sm = *reinterpret_cast<SweepSmem<KeyT>*>(smem_raw);
  typedef typename PairOf<KeyT>::type Pair;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (tid == 0) sm.tile = atomicAdd(ticket, 1u);
#pragma unroll
  for (int w = 0; w < kWarps; ++w) sm.warp_cnt[w][tid] = 0;
#pragma unroll
  for (int b = 0; b < 3; ++b)
#pragma unroll
    for (int w = 0; w < kWarps; ++w) sm.u.masks[b][w][tid] = 0;
  __syncthreads();
  const unsigned tile = sm.tile;
  const int64_t tile_base = (int64_t)tile * kTile;
  const bool full = tile_base + kTile <= n;

  // warp-striped load: warp w owns elements [w*512, (w+1)*512) of the tile, lane l takes element k*32 + l
  // in step k, so (warp, k, lane) order is input order.
  // Slots past the end of the input (last tile only) hold all-ones keys: they take the highest digit and, coming last
  // in input order, the last ranks of it, i.e. tile positions >= tile_n, which are never written out.  Nothing below
  // needs a validity test, and no tile looks back at the last one.
  KeyT key[kItems];
  uint32_t idx[kItems];
  const int slot0 = warp * (32 * kItems) + lane;
  const int64_t wbase = tile_base + slot0;
#pragma unroll
  for (int k = 0; k < kItems; ++k) key[k] = (full || wbase + k * 32 < n) ? __ldcs(keys_in + wbase + k * 32) : (KeyT)~(KeyT)0;
  // 64-bit keys leave no registers for the payload: it is fetched when the tile is reordered instead
  constexpr bool kLatePayload = sizeof(KeyT) == 8;
  if (!kLatePayload) {
#pragma unroll
    for (int k = 0; k < kItems; ++k) idx[k] = IOTA ? (uint32_t)(wbase + k * 32) : ((full || wbase + k * 32 < n) ? __ldcs(idx_in + wbase + k * 32) : 0u);
  }

  // rank inside the warp.  Lanes holding the same digit find each other by OR-ing their lane bit into a
  // per-(warp, digit) shared word; the lowest such lane advances the warp's running digit counter with one
  // shared-memory atomic.  The mask words are cleared by that leader one step later, hence three buffers and one
  // warp barrier per step.  (Shared-memory wavefronts bound this kernel: every access below is one per key.)
  unsigned rank2[kItems / 2];  // two 16-bit ranks per register
  const unsigned lt = (1u << lane) - 1u;
  unsigned* wc = sm.warp_cnt[warp];
  unsigned prev_d = 0;
  bool prev_leader = false;
#pragma unroll
  for (int k = 0; k < kItems; ++k) {
    unsigned* mk = sm.u.masks[k % 3][warp];
    const unsigned d = digit_of(key[k], shift, mask);
    atomicOr(&mk[d], 1u << lane);
    __syncwarp();
    if (k > 0 && prev_leader) sm.u.masks[(k + 2) % 3][warp][prev_d] = 0;
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
  __syncthreads();

  // per digit (thread tid == digit): the tile's count, published for the tiles behind; first slot in the tile;
  // each warp's first slot
  unsigned cnt[kWarps];
  unsigned run = 0;
#pragma unroll
  for (int w = 0; w < kWarps; ++w) {
    cnt[w] = sm.warp_cnt[w][tid];
    run += cnt[w];
  }
  LookT* mine = look + (int64_t)tile * kRadix + tid;
  Look<LookT>::store(mine, (LookT)run | ((LookT)(tile == 0 ? 2 : 1) << Look<LookT>::kShift));
  unsigned total;
  const unsigned bstart = block_exclusive_scan(run, sm.warp_s, &total);
  sm.bin_start[tid] = bstart;
  unsigned acc = bstart;
#pragma unroll
  for (int w = 0; w < kWarps; ++w) {
    sm.warp_cnt[w][tid] = acc;
    acc += cnt[w];
  }
  __syncthreads();

  // reorder inside shared memory (the mask words are dead: every warp passed the barriers above)
#pragma unroll
  for (int k = 0; k < kItems; ++k) {
    const unsigned d = digit_of(key[k], shift, mask);
    const unsigned pos = wc[d] + ((k & 1) ? (rank2[k / 2] >> 16) : (rank2[k / 2] & 0xffffu));
    Pair pr;
    pr.x = key[k];
    pr.y = !kLatePayload ? idx[k] : (IOTA ? (uint32_t)(wbase + k * 32) : ((full || wbase + k * 32 < n) ? __ldcs(idx_in + wbase + k * 32) : 0u));
    sm.u.out[pos] = pr;
  }

  // decoupled look-back: sum the counts of the tiles in front until one has published its inclusive count.
  // kLookWindow words are fetched at once (a hop is an L2 round trip); words behind the first inclusive one are
  // simply ignored, a word that is not ready yet is polled again.
  LookT excl = 0;
  if (tile > 0) {
    const LookT vmask = ((LookT)1 << Look<LookT>::kShift) - 1;
    int64_t t = (int64_t)tile - 1;
    bool done = false;
    while (!done) {
      LookT w[kLookWindow];
#pragma unroll
      for (int j = 0; j < kLookWindow; ++j) w[j] = (t - j >= 0) ? Look<LookT>::load(mine - (int64_t)(tile - (t - j)) * kRadix) : ((LookT)2 << Look<LookT>::kShift);
#pragma unroll
      for (int j = 0; j < kLookWindow; ++j) {
        if (done) break;
        LookT v = w[j];
        while ((v >> Look<LookT>::kShift) == 0) v = Look<LookT>::load(mine - (int64_t)(tile - (t - j)) * kRadix);
        excl += v & vmask;
        done = (v >> Look<LookT>::kShift) == 2;
      }
      t -= kLookWindow;
    }
    Look<LookT>::store(mine, (excl + run) | ((LookT)2 << Look<LookT>::kShift));
  }
  sm.delta[tid] = digit_base[tid] + (unsigned)excl - bstart;
  __syncthreads();

  // each digit's run goes out contiguously
  const int tile_n = (int)min((int64_t)kTile, n - tile_base);
#pragma unroll 4
  for (int p = tid; p < tile_n; p += kThreads) {
    const Pair pr = sm.u.out[p];
    const unsigned dst = (unsigned)p + sm.delta[digit_of((KeyT)pr.x, shift, mask)];
    keys_out[dst] = (KeyT)pr.x;
    idx_out[dst] = (uint32_t)pr.y;
  }
}

__global__ void iota_kernel(uint32_t* __restrict__ out, int64_t n) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) out[i] = (uint32_t)i;
}

inline int64_t num_tiles(int64_t n) { return (n + kTile - 1) / kTile; }
inline size_t align256(size_t v) { return (v + 255) / 256 * 256; }
// workspace: [digit histograms kMaxPasses x 256 u32][ticket (256 B)][look-back ntiles x 256 x 8 B]
inline size_t look_offset() { return align256(kMaxPasses * kRadix * sizeof(uint32_t)) + 256; }

bool g_force_wide_lookback = false;

template <typename KeyT, typename LookT>
int sweep_passes(KeyT* keys, KeyT* keys_alt, KeyT* keys_alt2, uint32_t* idx, uint32_t* idx_alt, int idx_is_iota, int64_t n, int end_bit,
                 int passes, void* ws, int32_t* result_location, cudaStream_t s) {
  const int64_t ntiles = num_tiles(n);
  unsigned char* base = static_cast<unsigned char*>(ws);
  uint32_t* ghist = reinterpret_cast<uint32_t*>(base);
  unsigned* ticket = reinterpret_cast<unsigned*>(base + align256(kMaxPasses * kRadix * sizeof(uint32_t)));
  LookT* look = reinterpret_cast<LookT*>(base + look_offset());
  const size_t look_bytes = (size_t)ntiles * kRadix * sizeof(LookT);
  const size_t smem = sizeof(SweepSmem<KeyT>);
  VAPC_RETURN_IF_CUDA(cudaFuncSetAttribute(onesweep_kernel<KeyT, LookT, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  VAPC_RETURN_IF_CUDA(cudaFuncSetAttribute(onesweep_kernel<KeyT, LookT, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  VAPC_RETURN_IF_CUDA(cudaMemsetAsync(ghist, 0, kMaxPasses * kRadix * sizeof(uint32_t), s));
  const int hgrid = (int)std::min<int64_t>((n + kThreads * 16 - 1) / (kThreads * 16), (int64_t)kNumSMs * 8);
  VAPC_LAUNCH(digit_hist_kernel<KeyT>, hgrid < 1 ? 1 : hgrid, kThreads, 0, s, keys, n, passes, end_bit, ghist);
  VAPC_LAUNCH(digit_scan_kernel, passes, kThreads, 0, s, ghist);
  // key buffers: keys -> alt -> (alt2 | keys) -> alt -> ...; with alt2 the caller's keys stay intact
  KeyT* src_k = keys; KeyT* dst_k = keys_alt;
  KeyT* spare = keys_alt2 ? keys_alt2 : keys;
  uint32_t* src_i = idx; uint32_t* dst_i = idx_alt;
  for (int p = 0; p < passes; ++p) {
    const int shift = p * kRadixBits;
    const int bits = std::min(kRadixBits, end_bit - shift);
    const unsigned mask = (1u << bits) - 1u;
    VAPC_RETURN_IF_CUDA(cudaMemsetAsync(ticket, 0, 256 + look_bytes, s));  // ticket + look-back words (contiguous)
    if (p == 0 && idx_is_iota)
      VAPC_LAUNCH((onesweep_kernel<KeyT, LookT, true>), (unsigned)ntiles, kThreads, smem, s, src_k, nullptr, dst_k, dst_i, n, shift, mask,
                  ghist + p * kRadix, look, ticket);
    else
      VAPC_LAUNCH((onesweep_kernel<KeyT, LookT, false>), (unsigned)ntiles, kThreads, smem, s, src_k, src_i, dst_k, dst_i, n, shift, mask,
                  ghist + p * kRadix, look, ticket);
    KeyT* next_dst = (p == 0) ? spare : src_k;
    src_k = dst_k; dst_k = next_dst;
    uint32_t* ti = src_i; src_i = dst_i; dst_i = ti;
  }
  VAPC_CHECK_LAUNCH();
  const int key_loc = src_k == keys ? 0 : (src_k == keys_alt ? 1 : 2);
  *result_location = key_loc | ((passes & 1) ? 4 : 0);
  return VAPC_OK;
}

template <typename KeyT>
int sort_impl(KeyT* keys, KeyT* keys_alt, KeyT* keys_alt2, uint32_t* idx, uint32_t* idx_alt, int idx_is_iota, int64_t n, int end_bit,
              void* ws, int32_t* result_location, cudaStream_t s) {
  const int passes = (end_bit + kRadixBits - 1) / kRadixBits;
  if (passes == 0) {
    if (idx_is_iota) VAPC_LAUNCH(iota_kernel, kNumSMs * 8, 256, 0, s, idx, n);
    *result_location = 0;
    VAPC_CHECK_LAUNCH();
    return VAPC_OK;
  }
  // 30-bit counts in 32-bit look-back words cover n < 2^30; beyond that (or when a test asks) 64-bit words
  if (n < ((int64_t)1 << 30) && !g_force_wide_lookback)
    return sweep_passes<KeyT, uint32_t>(keys, keys_alt, keys_alt2, idx, idx_alt, idx_is_iota, n, end_bit, passes, ws, result_location, s);
  return sweep_passes<KeyT, unsigned long long>(keys, keys_alt, keys_alt2, idx, idx_alt, idx_is_iota, n, end_bit, passes, ws, result_location, s);
}
}  // namespace

extern "C" int vapc_sort_debug_wide_lookback(int32_t on) {
  g_force_wide_lookback = on != 0;
  return VAPC_OK;
}

extern "C" size_t vapc_sort_workspace_bytes(int64_t n) {
  if (n < 1) n = 1;
  return look_offset() + (size_t)num_tiles(n) * kRadix * sizeof(unsigned long long);
}

extern "C" int vapc_sort_pairs(void* keys, void* keys_alt, void* keys_alt2, uint32_t* idx, uint32_t* idx_alt, int32_t idx_is_iota,
                               int64_t n, int32_t key_bytes, int32_t end_bit, void* ws, size_t ws_bytes, int32_t* result_location_host,
                               void* stream) {
  if (n < 0 || !result_location_host) return vapc_set_error(VAPC_E_INVALID, "vapc_sort_pairs: bad arguments");
  if (n >= (int64_t)1 << 32) return vapc_set_error(VAPC_E_INVALID, "vapc_sort_pairs: n must be < 2^32 per GPU");
  *result_location_host = 0;
  if (n == 0) return VAPC_OK;
  if (!keys || !keys_alt || !idx || !idx_alt) return vapc_set_error(VAPC_E_INVALID, "vapc_sort_pairs: null buffer");
  if (end_bit < 0 || end_bit > key_bytes * 8) return vapc_set_error(VAPC_E_INVALID, "vapc_sort_pairs: end_bit out of range");
  if (ws_bytes < vapc_sort_workspace_bytes(n) || !ws) return vapc_set_error(VAPC_E_WORKSPACE, "vapc_sort_pairs: workspace too small");
  if (key_bytes == 4)
    return sort_impl<uint32_t>(static_cast<uint32_t*>(keys), static_cast<uint32_t*>(keys_alt), static_cast<uint32_t*>(keys_alt2), idx, idx_alt,
                               idx_is_iota, n, end_bit, ws, result_location_host, as_stream(stream));
  if (key_bytes == 8)
    return sort_impl<unsigned long long>(static_cast<unsigned long long*>(keys), static_cast<unsigned long long*>(keys_alt),
                                         static_cast<unsigned long long*>(keys_alt2), idx, idx_alt, idx_is_iota, n, end_bit, ws,
                                         result_location_host, as_stream(stream));
  return vapc_set_error(VAPC_E_INVALID, "vapc_sort_pairs: key_bytes must be 4 or 8");
}
