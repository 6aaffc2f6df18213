// In-house stable LSD radix sort of (key, point index) pairs — 8-bit digits, ONE sweep per pass.
//
//   digit_hist_kernel   one read of the keys -> the digit histogram of EVERY pass (shared-memory atomics);
//                       digit_scan_kernel turns each into first-slot offsets                          K B/elt
//   onesweep_kernel     per pass; tile = 4096 consecutive elements, one CTA of 256 threads: rank inside the
//                       tile (radix_rank.cuh), publish the tile's digit counts, decoupled look-back over the
//                       tiles in front for the global offsets, reorder the tile in shared memory and write
//                       each digit's run coalesced                    reads K+4, writes K+4 B/elt, nothing else
// Ranking keeps equal digits in input order, so every pass — and therefore the whole sort — is stable:
// points of one voxel stay in original index order (tie-break "lowest original index" and "first point of a
// voxel" fall out of that; see include/vapc_b200.h).
//
// Segmented form (vapc_sort_pairs_segmented): the array is a sequence of <= 256 segments (the key-prefix
// buckets of vapc_pack_records_bucketed) that are sorted independently by their low key bits in the same
// launches: tiles never straddle a segment, histograms and look-back chains are per segment.
#include <algorithm>

#include "radix_rank.cuh"

namespace {
using namespace rr;
#ifndef VAPC_SORT_ITEMS
#define VAPC_SORT_ITEMS 16
#endif
#ifndef VAPC_SORT_MINB
#define VAPC_SORT_MINB 4
#endif
constexpr int kItems = VAPC_SORT_ITEMS;
constexpr int kTile = kThreads * kItems;  // 4096
constexpr int kMaxPasses = 8;             // 64-bit keys
constexpr int kMaxSegments = 256;

template <typename KeyT>
__device__ __forceinline__ unsigned digit_of(KeyT k, int shift, unsigned mask) {
  return (unsigned)(k >> shift) & mask;
}

// ---- tiles of a segmented array -------------------------------------------------------------------------
// seg_first_tile[s] = first tile of segment s (exclusive scan of ceil(len / kTile)); [nseg] = number of tiles
__global__ void __launch_bounds__(kThreads) seg_tiles_kernel(const uint32_t* __restrict__ seg_start, int nseg, uint32_t* __restrict__ seg_first_tile) {
  __shared__ unsigned warp_s[33];
  const int s = threadIdx.x;
  const unsigned len = s < nseg ? seg_start[s + 1] - seg_start[s] : 0u;
  const unsigned tiles = (len + kTile - 1) / kTile;
  unsigned total;
  const unsigned first = block_exclusive_scan(tiles, warp_s, &total);
  if (s < nseg) seg_first_tile[s] = first;
  if (s == 0) seg_first_tile[nseg] = total;
}
// segment of tile t: the last s with first[s] <= t (empty segments have first[s] == first[s+1] and are skipped)
__device__ __forceinline__ int segment_of_tile(const unsigned* first, int nseg, unsigned t) {
  int lo = 0, hi = nseg;  // invariant: first[lo] <= t < first[hi]
  while (hi - lo > 1) {
    const int mid = (lo + hi) >> 1;
    if (first[mid] <= t) lo = mid; else hi = mid;
  }
  return lo;
}

// ---- all digit histograms in one sweep over the keys ---------------------------------------------
// ghist[p][s][d] = number of keys of segment s whose p-th 8-bit digit is d.  Shared-memory atomics (no
// return value needed for counting); a CTA walks a contiguous range of tiles and flushes its non-zero
// counters when the segment changes.
// PASSES > 0: the number of passes as a compile-time constant (the kernel is bound by issue slots, not by the shared-memory
// atomics: with a run-time pass loop it spent 47 instructions per key and warp); 0: any number of passes.
template <typename KeyT, bool SEG, int PASSES>
__global__ void __launch_bounds__(kThreads) digit_hist_kernel(const KeyT* __restrict__ keys, int64_t n, int passes_rt, int begin_bit, int end_bit,
                                                              uint32_t* __restrict__ ghist, const uint32_t* __restrict__ seg_start,
                                                              const uint32_t* __restrict__ seg_first_tile, int nseg, int tiles_per_cta) {
  __shared__ unsigned sh[PASSES ? PASSES : kMaxPasses][kRadix];
  __shared__ unsigned first[kMaxSegments + 1];
  const int passes = PASSES ? PASSES : passes_rt;
#pragma unroll
  for (int p = 0; p < passes; ++p) sh[p][threadIdx.x] = 0;
  if (SEG) {
    if (threadIdx.x <= nseg) first[threadIdx.x] = seg_first_tile[threadIdx.x];
    if (threadIdx.x == 0 && nseg == kMaxSegments) first[kMaxSegments] = seg_first_tile[kMaxSegments];
  }
  __syncthreads();
  const unsigned last_mask = (1u << (end_bit - begin_bit - (passes - 1) * kRadixBits)) - 1u;
  const int64_t ntiles = SEG ? (int64_t)first[nseg] : (n + kTile - 1) / kTile;
  const int64_t t0 = (int64_t)blockIdx.x * tiles_per_cta;
  const int64_t t1 = min(ntiles, t0 + tiles_per_cta);
  int cur = -1;
  for (int64_t t = t0; t < t1; ++t) {
    int64_t base, end;
    if (SEG) {
      const int s = segment_of_tile(first, nseg, (unsigned)t);
      if (s != cur) {
        if (cur >= 0) {
          __syncthreads();
          for (int p = 0; p < passes; ++p) {
            const unsigned c = sh[p][threadIdx.x];
            if (c) atomicAdd(&ghist[((int64_t)p * nseg + cur) * kRadix + threadIdx.x], c);
            sh[p][threadIdx.x] = 0;
          }
          __syncthreads();
        }
        cur = s;
      }
      base = (int64_t)seg_start[s] + (t - first[s]) * kTile;
      end = min((int64_t)seg_start[s + 1], base + kTile);
    } else {
      cur = 0;
      base = t * kTile;
      end = min(n, base + kTile);
    }
    // a tile is at most kItems keys per thread: all loads of the tile are issued before the first atomic
    KeyT k[kItems];
#pragma unroll
    for (int j = 0; j < kItems; ++j) {
      const int64_t i = base + j * kThreads + threadIdx.x;
      k[j] = i < end ? __ldcs(keys + i) : (KeyT)0;
    }
    if (end - base == kTile) {  // a full tile: no per-key bound test
#pragma unroll
      for (int j = 0; j < kItems; ++j) {
#pragma unroll
        for (int p = 0; p < passes; ++p) {
          const unsigned m = p == passes - 1 ? last_mask : (unsigned)(kRadix - 1);
          atomicAdd(&sh[p][digit_of(k[j], begin_bit + p * kRadixBits, m)], 1u);
        }
      }
    } else {
#pragma unroll
      for (int j = 0; j < kItems; ++j) {
        if (base + j * kThreads + threadIdx.x < end) {
#pragma unroll
          for (int p = 0; p < passes; ++p) {
            const unsigned m = p == passes - 1 ? last_mask : (unsigned)(kRadix - 1);
            atomicAdd(&sh[p][digit_of(k[j], begin_bit + p * kRadixBits, m)], 1u);
          }
        }
      }
    }
  }
  __syncthreads();
  if (cur >= 0) {
    for (int p = 0; p < passes; ++p) {
      const unsigned c = sh[p][threadIdx.x];
      if (c) atomicAdd(&ghist[((int64_t)p * nseg + cur) * kRadix + threadIdx.x], c);
    }
  }
}

// ---- one radix pass in one sweep: rank + decoupled look-back + scatter --------------------------------
template <typename KeyT> struct PairOf;  // (key, payload) as stored in the reorder buffer
template <> struct PairOf<uint32_t> { typedef uint2 type; };           // one 64-bit shared-memory access per element
template <> struct PairOf<unsigned long long> { typedef ulonglong2 type; };  // key + zero-extended payload

template <typename KeyT>
struct SweepSmem {
  RankSmem r;
  unsigned seg_first[kMaxSegments + 1];
  alignas(16) union {
    MaskBuffers masks;
    typename PairOf<KeyT>::type out[kTile];  // the tile reordered by digit
  } u;
};

template <typename KeyT, typename LookT, bool IOTA, bool SEG>
__global__ void __launch_bounds__(kThreads, sizeof(KeyT) == 4 ? VAPC_SORT_MINB : 3) onesweep_kernel(
    const KeyT* __restrict__ keys_in, const uint32_t* __restrict__ idx_in, KeyT* __restrict__ keys_out, uint32_t* __restrict__ idx_out,
    int64_t n, int shift, unsigned mask, const uint32_t* __restrict__ digit_base, LookT* __restrict__ look, unsigned* __restrict__ ticket,
    const uint32_t* __restrict__ seg_start, const uint32_t* __restrict__ seg_first_tile, int nseg) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  SweepSmem<KeyT>& sm = *reinterpret_cast<SweepSmem<KeyT>*>(smem_raw);
  typedef typename PairOf<KeyT>::type Pair;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  // Tiles are handed out by an atomic ticket, so a tile's predecessors are always resident or done.
  if (tid == 0) sm.r.tile = atomicAdd(ticket, 1u);
  rank_init(sm.r, sm.u.masks);
  if (SEG) {
    if (tid <= nseg) sm.seg_first[tid] = seg_first_tile[tid];
    if (tid == 0 && nseg == kMaxSegments) sm.seg_first[kMaxSegments] = seg_first_tile[kMaxSegments];
  }
  __syncthreads();
  const unsigned tile = sm.r.tile;
  int64_t tile_base, tile_end, in_front;
  if (SEG) {
    if (tile >= sm.seg_first[nseg]) return;  // the grid is sized for the worst case
    const int s = segment_of_tile(sm.seg_first, nseg, tile);
    in_front = (int64_t)tile - sm.seg_first[s];
    tile_base = (int64_t)seg_start[s] + in_front * kTile;
    tile_end = min((int64_t)seg_start[s + 1], tile_base + kTile);
    digit_base += (int64_t)s * kRadix;
  } else {
    in_front = tile;
    tile_base = (int64_t)tile * kTile;
    tile_end = min(n, tile_base + kTile);
  }
  const bool full = tile_end - tile_base == kTile;

  // warp-striped load: warp w owns elements [w*512, (w+1)*512) of the tile, lane l takes element k*32 + l
  // in step k, so (warp, k, lane) order is input order.
  // Slots past the end of the tile (last tile of the array / of a segment) hold all-ones keys: they take the
  // highest digit and, coming last in input order, the last ranks of it, i.e. tile positions >= the tile's
  // length, which are never written out.  Nothing below needs a validity test, and no tile looks back at a
  // partial one.
  KeyT key[kItems];
  uint32_t idx[kItems];
  const int64_t wbase = tile_base + warp * (32 * kItems) + lane;
#pragma unroll
  for (int k = 0; k < kItems; ++k) key[k] = (full || wbase + k * 32 < tile_end) ? __ldcs(keys_in + wbase + k * 32) : (KeyT)~(KeyT)0;
  // 64-bit keys leave no registers for the payload: it is fetched when the tile is reordered instead
  constexpr bool kLatePayload = sizeof(KeyT) == 8;
  if (!kLatePayload) {
#pragma unroll
    for (int k = 0; k < kItems; ++k)
      idx[k] = IOTA ? (uint32_t)(wbase + k * 32) : ((full || wbase + k * 32 < tile_end) ? __ldcs(idx_in + wbase + k * 32) : 0u);
  }

  unsigned rank2[kItems / 2];
  warp_rank<kItems>([&](int k) { return digit_of(key[k], shift, mask); }, sm.r, sm.u.masks, rank2);
  __syncthreads();

  LookT* mine = look + (int64_t)tile * kRadix + tid;
  unsigned bstart;
  const unsigned run = digit_offsets<LookT>(sm.r, mine, in_front == 0, &bstart);

  // reorder inside shared memory (the mask words are dead: every warp passed the barriers above)
  const unsigned* wc = sm.r.warp_cnt[warp];
#pragma unroll
  for (int k = 0; k < kItems; ++k) {
    const unsigned pos = wc[digit_of(key[k], shift, mask)] + rank_of(rank2, k);
    Pair pr;
    pr.x = key[k];
    pr.y = !kLatePayload ? idx[k] : (IOTA ? (uint32_t)(wbase + k * 32) : ((full || wbase + k * 32 < tile_end) ? __ldcs(idx_in + wbase + k * 32) : 0u));
    sm.u.out[pos] = pr;
  }

  const LookT excl = look_back<LookT>(mine, in_front, run);
  sm.r.delta[tid] = digit_base[tid] + (unsigned)excl - bstart;
  __syncthreads();

  // each digit's run goes out contiguously
  const int tile_n = (int)(tile_end - tile_base);
#pragma unroll 4
  for (int p = tid; p < tile_n; p += kThreads) {
    const Pair pr = sm.u.out[p];
    const unsigned dst = (unsigned)p + sm.r.delta[digit_of((KeyT)pr.x, shift, mask)];
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
// workspace: [digit histograms kMaxPasses x kMaxSegments x 256 u32][seg_first_tile][ticket (256 B)][look-back tiles x 256 x 8 B]
inline size_t hist_bytes() { return (size_t)kMaxPasses * kMaxSegments * kRadix * sizeof(uint32_t); }
inline size_t segtab_bytes() { return align256((kMaxSegments + 1) * sizeof(uint32_t)); }
inline size_t look_offset() { return hist_bytes() + segtab_bytes() + 256; }

bool g_force_wide_lookback = false;

template <typename KeyT, typename LookT, bool SEG>
int sweep_passes(KeyT* keys, KeyT* keys_alt, KeyT* keys_alt2, uint32_t* idx, uint32_t* idx_alt, int idx_is_iota, int64_t n, int begin_bit,
                 int end_bit, int passes, const uint32_t* seg_start, int nseg, void* ws, int32_t* result_location, cudaStream_t s) {
  const int64_t ntiles = num_tiles(n) + (SEG ? nseg : 0);  // segmented: upper bound, surplus CTAs exit
  unsigned char* base = static_cast<unsigned char*>(ws);
  uint32_t* ghist = reinterpret_cast<uint32_t*>(base);
  uint32_t* seg_first_tile = reinterpret_cast<uint32_t*>(base + hist_bytes());
  unsigned* ticket = reinterpret_cast<unsigned*>(base + hist_bytes() + segtab_bytes());
  LookT* look = reinterpret_cast<LookT*>(base + look_offset());
  const size_t look_bytes = (size_t)ntiles * kRadix * sizeof(LookT);
  const size_t smem = sizeof(SweepSmem<KeyT>);
  VAPC_RETURN_IF_CUDA(cudaFuncSetAttribute(onesweep_kernel<KeyT, LookT, true, SEG>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  VAPC_RETURN_IF_CUDA(cudaFuncSetAttribute(onesweep_kernel<KeyT, LookT, false, SEG>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const int hseg = SEG ? nseg : 1;
  VAPC_RETURN_IF_CUDA(cudaMemsetAsync(ghist, 0, (size_t)passes * hseg * kRadix * sizeof(uint32_t), s));
  if (SEG) VAPC_LAUNCH(seg_tiles_kernel, 1, kThreads, 0, s, seg_start, nseg, seg_first_tile);
  const int hgrid = (int)std::max<int64_t>(1, std::min<int64_t>(ntiles, (int64_t)kNumSMs * 8));
  const int tiles_per_cta = (int)((ntiles + hgrid - 1) / hgrid);
#define VAPC_HIST(P) VAPC_LAUNCH((digit_hist_kernel<KeyT, SEG, P>), hgrid, kThreads, 0, s, keys, n, passes, begin_bit, end_bit, ghist, seg_start, seg_first_tile, hseg, tiles_per_cta)
  switch (passes) {
    case 1: VAPC_HIST(1); break;
    case 2: VAPC_HIST(2); break;
    case 3: VAPC_HIST(3); break;
    case 4: VAPC_HIST(4); break;
    default: VAPC_HIST(0); break;
  }
#undef VAPC_HIST
  VAPC_LAUNCH(digit_scan_kernel, passes * hseg, kThreads, 0, s, ghist, SEG ? seg_start : nullptr, hseg);
  // key buffers: keys -> alt -> (alt2 | keys) -> alt -> ...; with alt2 the caller's keys stay intact
  KeyT* src_k = keys; KeyT* dst_k = keys_alt;
  KeyT* spare = keys_alt2 ? keys_alt2 : keys;
  uint32_t* src_i = idx; uint32_t* dst_i = idx_alt;
  for (int p = 0; p < passes; ++p) {
    const int shift = begin_bit + p * kRadixBits;
    const int bits = std::min(kRadixBits, end_bit - shift);
    const unsigned mask = (1u << bits) - 1u;
    VAPC_RETURN_IF_CUDA(cudaMemsetAsync(ticket, 0, 256 + look_bytes, s));  // ticket + look-back words (contiguous)
    const uint32_t* dbase = ghist + (int64_t)p * hseg * kRadix;
    if (p == 0 && idx_is_iota)
      VAPC_LAUNCH((onesweep_kernel<KeyT, LookT, true, SEG>), (unsigned)ntiles, kThreads, smem, s, src_k, nullptr, dst_k, dst_i, n, shift, mask, dbase,
                  look, ticket, seg_start, seg_first_tile, hseg);
    else
      VAPC_LAUNCH((onesweep_kernel<KeyT, LookT, false, SEG>), (unsigned)ntiles, kThreads, smem, s, src_k, src_i, dst_k, dst_i, n, shift, mask, dbase,
                  look, ticket, seg_start, seg_first_tile, hseg);
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
int sort_impl(KeyT* keys, KeyT* keys_alt, KeyT* keys_alt2, uint32_t* idx, uint32_t* idx_alt, int idx_is_iota, int64_t n, int begin_bit,
              int end_bit, const uint32_t* seg_start, int nseg, void* ws, int32_t* result_location, cudaStream_t s) {
  const int passes = (end_bit - begin_bit + kRadixBits - 1) / kRadixBits;
  if (passes == 0) {
    if (idx_is_iota) VAPC_LAUNCH(iota_kernel, kNumSMs * 8, 256, 0, s, idx, n);
    *result_location = 0;
    VAPC_CHECK_LAUNCH();
    return VAPC_OK;
  }
  // 30-bit counts in 32-bit look-back words cover n < 2^30; beyond that (or when a test asks) 64-bit words
  const bool wide = n >= ((int64_t)1 << 30) || g_force_wide_lookback;
  if (seg_start) {
    if (!wide) return sweep_passes<KeyT, uint32_t, true>(keys, keys_alt, keys_alt2, idx, idx_alt, idx_is_iota, n, begin_bit, end_bit, passes, seg_start, nseg, ws, result_location, s);
    return sweep_passes<KeyT, unsigned long long, true>(keys, keys_alt, keys_alt2, idx, idx_alt, idx_is_iota, n, begin_bit, end_bit, passes, seg_start, nseg, ws, result_location, s);
  }
  if (!wide) return sweep_passes<KeyT, uint32_t, false>(keys, keys_alt, keys_alt2, idx, idx_alt, idx_is_iota, n, begin_bit, end_bit, passes, nullptr, 1, ws, result_location, s);
  return sweep_passes<KeyT, unsigned long long, false>(keys, keys_alt, keys_alt2, idx, idx_alt, idx_is_iota, n, begin_bit, end_bit, passes, nullptr, 1, ws, result_location, s);
}

int sort_entry(void* keys, void* keys_alt, void* keys_alt2, uint32_t* idx, uint32_t* idx_alt, int32_t idx_is_iota, int64_t n, int32_t key_bytes,
               int32_t begin_bit, int32_t end_bit, const uint32_t* seg_start, int32_t nseg, void* ws, size_t ws_bytes, int32_t* result_location_host, void* stream,
               const char* who) {
  if (n < 0 || !result_location_host) return vapc_set_error(VAPC_E_INVALID, "%s: bad arguments", who);
  if (n >= (int64_t)1 << 32) return vapc_set_error(VAPC_E_INVALID, "%s: n must be < 2^32 per GPU", who);
  *result_location_host = 0;
  if (n == 0) return VAPC_OK;
  if (!keys || !keys_alt || !idx || !idx_alt) return vapc_set_error(VAPC_E_INVALID, "%s: null buffer", who);
  if (begin_bit < 0 || end_bit < begin_bit || end_bit > key_bytes * 8) return vapc_set_error(VAPC_E_INVALID, "%s: bit range out of bounds", who);
  if (ws_bytes < vapc_sort_workspace_bytes(n) || !ws) return vapc_set_error(VAPC_E_WORKSPACE, "%s: workspace too small", who);
  if (key_bytes == 4)
    return sort_impl<uint32_t>(static_cast<uint32_t*>(keys), static_cast<uint32_t*>(keys_alt), static_cast<uint32_t*>(keys_alt2), idx, idx_alt,
                               idx_is_iota, n, begin_bit, end_bit, seg_start, nseg, ws, result_location_host, as_stream(stream));
  if (key_bytes == 8)
    return sort_impl<unsigned long long>(static_cast<unsigned long long*>(keys), static_cast<unsigned long long*>(keys_alt),
                                         static_cast<unsigned long long*>(keys_alt2), idx, idx_alt, idx_is_iota, n, begin_bit, end_bit, seg_start, nseg, ws,
                                         result_location_host, as_stream(stream));
  return vapc_set_error(VAPC_E_INVALID, "%s: key_bytes must be 4 or 8", who);
}
}  // namespace

extern "C" int vapc_sort_debug_wide_lookback(int32_t on) {
  g_force_wide_lookback = on != 0;
  return VAPC_OK;
}

extern "C" size_t vapc_sort_workspace_bytes(int64_t n) {
  if (n < 1) n = 1;
  return look_offset() + (size_t)(num_tiles(n) + kMaxSegments) * kRadix * sizeof(unsigned long long);
}

extern "C" int vapc_sort_pairs(void* keys, void* keys_alt, void* keys_alt2, uint32_t* idx, uint32_t* idx_alt, int32_t idx_is_iota,
                               int64_t n, int32_t key_bytes, int32_t end_bit, void* ws, size_t ws_bytes, int32_t* result_location_host,
                               void* stream) {
  return sort_entry(keys, keys_alt, keys_alt2, idx, idx_alt, idx_is_iota, n, key_bytes, 0, end_bit, nullptr, 1, ws, ws_bytes, result_location_host,
                    stream, "vapc_sort_pairs");
}

extern "C" int vapc_sort_pairs_bits(void* keys, void* keys_alt, void* keys_alt2, uint32_t* idx, uint32_t* idx_alt, int32_t idx_is_iota,
                                    int64_t n, int32_t key_bytes, int32_t begin_bit, int32_t end_bit, void* ws, size_t ws_bytes,
                                    int32_t* result_location_host, void* stream) {
  return sort_entry(keys, keys_alt, keys_alt2, idx, idx_alt, idx_is_iota, n, key_bytes, begin_bit, end_bit, nullptr, 1, ws, ws_bytes,
                    result_location_host, stream, "vapc_sort_pairs_bits");
}

extern "C" int vapc_sort_pairs_segmented(void* keys, void* keys_alt, uint32_t* idx, uint32_t* idx_alt, int32_t idx_is_iota, int64_t n,
                                         int32_t key_bytes, int32_t end_bit, const uint32_t* seg_start, int32_t nseg, void* ws, size_t ws_bytes,
                                         int32_t* result_location_host, void* stream) {
  if (!seg_start || nseg < 1 || nseg > kMaxSegments)
    return vapc_set_error(VAPC_E_INVALID, "vapc_sort_pairs_segmented: 1..%d segments", kMaxSegments);
  return sort_entry(keys, keys_alt, nullptr, idx, idx_alt, idx_is_iota, n, key_bytes, 0, end_bit, seg_start, nseg, ws, ws_bytes, result_location_host,
                    stream, "vapc_sort_pairs_segmented");
}
