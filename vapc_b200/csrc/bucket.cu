// Keys + 32-byte point records, written in KEY-PREFIX BUCKET order instead of input order.
//
// Why: the per-voxel reduction gathers one record per point in sorted-key order.  With the records in input
// order those are random 32-byte reads over the whole array and every one of them costs a full 128-byte line
// from HBM (measured: 13.4 GB of DRAM reads for 3.2 GB of records at 1e8 points, profiles/).  Here the
// records are grouped by the leading bits of the packed key — the high bits of the voxel x index, <= 256
// buckets — while they are produced, so that the later gathers of one bucket stay inside a window of
// records that fits the 126 MB L2 and every line is fetched from HBM once.  The buckets are also exactly the
// segments that vapc_sort_pairs_segmented sorts by the REMAINING key bits, which saves the radix pass over
// the leading digit.
//
// Round 2 layout (no ticket, no decoupled look-back: 25 % of the stall samples of the round-1 scatter kernel were its poll):
//   tile_digits_kernel    a tile = 1536 consecutive points.  Reads x ONLY (the bucket digit is the leading bits of the voxel x
//                         index) and writes the tile's 256 digit counts                                  8 B/pt in, 0.33 B/pt out
//   tile_chunk_sums / tile_chunk_scan / tile_offsets
//                         column scan of the [tiles][256] count matrix -> first output slot of every (tile, digit),
//                         bucket_start[]                                                                  ~1.3 B/pt
//   tile_scatter_kernel   reads x, y, z, forms the packed key (floor((c - o) / vs) per axis, bit-exact without the division for almost
//                         every point: voxel_of_fast), takes the digit from the voxel x index, ranks the digits (radix_rank.cuh,
//                         stable), writes the key in input order, and key + record {x, y, z, index} at the bucketed position as
//                         whole lines                                                                     24.7 B/pt in, 40 B/pt out
// (Round 1: keys first in a pass of their own over x, y, z (28 B/pt), then a scatter pass with tickets and look-back (64 B/pt);
// a kernel that published its counts only after the three divisions had spent 34 % of its samples polling predecessors.  A digit
// byte per point written by the counting pass and read back by the scatter pass — VAPC_BUCKET_DIGIT_BYTES=1 — cost 0.1 ms more.)
// Stable: inside a bucket the points keep their input order, so after the (stable) segmented sort the points
// of a voxel are still in original index order.
#include <algorithm>

#include "radix_rank.cuh"

namespace {
using namespace rr;
#ifndef VAPC_BUCKET_ITEMS
#define VAPC_BUCKET_ITEMS 6
#endif
#ifndef VAPC_BUCKET_MINB
#define VAPC_BUCKET_MINB 3
#endif
#ifndef VAPC_BUCKET_DIGIT_BYTES
#define VAPC_BUCKET_DIGIT_BYTES 0  // 1: the counting pass also writes a digit byte per point that the scatter pass reads back (2 B/point more traffic;
                                   // the scatter kernel forms the voxel x index for the key anyway, so it takes the digit from there)
#endif
constexpr bool kDigitBytes = VAPC_BUCKET_DIGIT_BYTES != 0;
constexpr int kItems = VAPC_BUCKET_ITEMS;
constexpr int kTile = kThreads * kItems;  // 1536 points: the reordered records of a tile take 48 KB of shared memory (6 vs 8 items: 2.22 vs 2.42 ms)
constexpr int kChunkTiles = 128;          // tiles per chunk of the column scan (512: 128 CTAs, each a serial loop of 512 loads — 0.04 ms per kernel)
constexpr int kScanGroups = 4;            // the chunk scan runs 4 column scans side by side, each over a quarter of the chunks

template <typename KeyT>
__device__ __forceinline__ unsigned bucket_of(KeyT key, int shift, unsigned mask) { return (unsigned)(key >> shift) & mask; }

// digit byte of every point + the digit counts of every tile.  dshift = bits of the voxel x index below the bucket digit.
template <typename Src>
__global__ void __launch_bounds__(kThreads) tile_digits_kernel(Src src, int64_t n, Grid g, int dshift, unsigned mask, uint8_t* __restrict__ digits,
                                                               uint16_t* __restrict__ tile_counts, int* __restrict__ status) {
  __shared__ unsigned sh[kRadix];
  sh[threadIdx.x] = 0;
  __syncthreads();
  const int64_t base = (int64_t)blockIdx.x * kTile + (threadIdx.x >> 5) * (32 * kItems) + (threadIdx.x & 31);
  const bool full = (int64_t)(blockIdx.x + 1) * kTile <= n;  // all but the last tile: no per-point bound test
  double px[kItems];
#pragma unroll
  for (int k = 0; k < kItems; ++k) px[k] = (full || base + k * 32 < n) ? src.cx(base + k * 32) : 0.0;
  unsigned long long over = 0;  // bits of any relative x index beyond the grid's bx bits
#pragma unroll
  for (int k = 0; k < kItems; ++k) {
    if (full || base + k * 32 < n) {
      const unsigned long long rx = (unsigned long long)(voxel_of_fast(px[k], g.ox, g.vs, g.ivs) - g.minx);
      over |= rx;
      const unsigned d = (unsigned)(rx >> dshift) & mask;  // an out-of-grid point gets SOME digit: the caller raises on the status bit
      if (kDigitBytes) digits[base + k * 32] = (uint8_t)d;
      atomicAdd(&sh[d], 1u);
    }
  }
  const int bad = g.bx < 64 && (over >> g.bx) != 0;
  __syncthreads();
  tile_counts[(int64_t)blockIdx.x * kRadix + threadIdx.x] = (uint16_t)sh[threadIdx.x];
  if (__any_sync(0xffffffffu, bad) && (threadIdx.x & 31) == 0) atomicOr(status, 1);
}

// column scan of tile_counts[tiles][256], thread == digit: chunk sums -> exclusive over chunks + digits -> per-tile offsets
__global__ void __launch_bounds__(kThreads) tile_chunk_sums_kernel(const uint16_t* __restrict__ tile_counts, int64_t ntiles, uint32_t* __restrict__ chunk_sums) {
  const int64_t t0 = (int64_t)blockIdx.x * kChunkTiles, t1 = min(ntiles, t0 + kChunkTiles);
  unsigned sum = 0;
#pragma unroll 16
  for (int64_t t = t0; t < t1; ++t) sum += tile_counts[t * kRadix + threadIdx.x];
  chunk_sums[(int64_t)blockIdx.x * kRadix + threadIdx.x] = sum;
}
// chunk_sums[c][d] -> exclusive prefix over the chunks of the thread's GROUP (a quarter of the chunks, scanned side by side by the four
// groups of the CTA); group_add[g][d] = first slot of digit d's bucket + the keys of d in the groups in front of g.  The first slot of
// (chunk c, digit d) is chunk_sums[c][d] + group_add[c / per][d]: tile_offsets_kernel adds the two.  Loads are issued eight chunks
// ahead of the stores (a load - store - load chain through one pointer is a DRAM round trip per chunk).
__global__ void __launch_bounds__(kThreads * kScanGroups) tile_chunk_scan_kernel(uint32_t* __restrict__ chunk_sums, int nchunks, int per, int nb, uint32_t n,
                                                                                 uint32_t* __restrict__ group_add, uint32_t* __restrict__ bucket_start) {
  __shared__ unsigned warp_s[33];
  __shared__ unsigned group_total[kScanGroups][kRadix];
  const int d = threadIdx.x & (kRadix - 1), grp = threadIdx.x >> kRadixBits;
  const int c0 = min(nchunks, grp * per), c1 = min(nchunks, c0 + per);
  unsigned total = 0;
  for (int c = c0; c < c1; c += 8) {
    unsigned v[8];
#pragma unroll
    for (int q = 0; q < 8; ++q) v[q] = c + q < c1 ? chunk_sums[(int64_t)(c + q) * kRadix + d] : 0u;
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      if (c + q < c1) chunk_sums[(int64_t)(c + q) * kRadix + d] = total;
      total += v[q];
    }
  }
  group_total[grp][d] = total;
  __syncthreads();
  unsigned before = 0, all_groups = 0;  // keys of digit d in the groups in front / in all chunks
#pragma unroll
  for (int q = 0; q < kScanGroups; ++q) {
    const unsigned t = group_total[q][d];
    if (q < grp) before += t;
    all_groups += t;
  }
  // first slot of the digit's bucket: exclusive scan over the digits (the threads of the other groups add zeros)
  __shared__ unsigned digit_base[kRadix];
  unsigned all;
  const unsigned base = block_exclusive_scan(grp == 0 ? all_groups : 0u, warp_s, &all);
  if (grp == 0) digit_base[d] = base;
  __syncthreads();
  group_add[grp * kRadix + d] = digit_base[d] + before;
  if (grp == 0) {
    if (d < nb) bucket_start[d] = digit_base[d];
    if (d == 0) bucket_start[nb] = n;
  }
}
__global__ void __launch_bounds__(kThreads) tile_offsets_kernel(const uint16_t* __restrict__ tile_counts, const uint32_t* __restrict__ chunk_base,
                                                                const uint32_t* __restrict__ group_add, int per, int64_t ntiles,
                                                                uint32_t* __restrict__ tile_offsets) {
  const int64_t t0 = (int64_t)blockIdx.x * kChunkTiles, t1 = min(ntiles, t0 + kChunkTiles);
  unsigned run = chunk_base[(int64_t)blockIdx.x * kRadix + threadIdx.x] + group_add[(blockIdx.x / per) * kRadix + threadIdx.x];
#pragma unroll 16
  for (int64_t t = t0; t < t1; ++t) {
    tile_offsets[t * kRadix + threadIdx.x] = run;
    run += tile_counts[t * kRadix + threadIdx.x];
  }
}

// The tile's records reordered by bucket, as columns: coordinates, key and the 16-bit input slot of every point (64-bit
// shared-memory stores per column measured 4 % faster than one 32-byte record store; a fourth CTA per SM — 55 KB each —
// needs 64 registers per thread and loses more to spills than it gains).
template <typename KeyT>
struct BucketSmem {
  RankSmem r;
  alignas(32) union {
    MaskBuffers masks;
    struct {
      double x[kTile], y[kTile], z[kTile];
      KeyT key[kTile];
      unsigned short slot[kTile];
    } rec;
  } u;
};

// KeyOutT: the bucketed keys are written as 32-bit words holding the key bits BELOW the bucket digit when those fit
// (64-bit keys of up to 40 bits): the segmented sort then moves 4-byte keys; vapc_expand_bucket_keys restores the full keys.
template <typename KeyT, typename KeyOutT, typename Src>
__global__ void __launch_bounds__(kThreads, sizeof(KeyT) == 4 ? VAPC_BUCKET_MINB : 3) tile_scatter_kernel(
    const uint8_t* __restrict__ digits, Src src, int64_t n, Grid g, int shift, int dshift, unsigned mask, const uint32_t* __restrict__ tile_offsets,
    KeyT* __restrict__ keys_orig, KeyOutT* __restrict__ keys_b, double4* __restrict__ rec_b, int* __restrict__ status) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  BucketSmem<KeyT>& sm = *reinterpret_cast<BucketSmem<KeyT>*>(smem_raw);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  rank_init(sm.r, sm.u.masks);
  const unsigned tile = blockIdx.x;
  const int64_t tile_base = (int64_t)tile * kTile;
  const bool full = tile_base + kTile <= n;
  const unsigned my_offset = tile_offsets[(int64_t)tile * kRadix + tid];  // thread == digit: first output slot of the tile's run
  // warp-striped: warp w owns points [w*32*kItems, (w+1)*32*kItems) of the tile, lane l takes point k*32 + l in step k.
  // Slots past the end of the input (last tile) take the highest digit: last ranks of it, never written.
  const int slot0 = warp * (32 * kItems) + lane;
  const int64_t wbase = tile_base + slot0;
  unsigned dig[kItems];
  if (kDigitBytes) {
#pragma unroll
    for (int k = 0; k < kItems; ++k) dig[k] = (full || wbase + k * 32 < n) ? (unsigned)digits[wbase + k * 32] : (unsigned)(kRadix - 1);
  }
  double px[kItems], py[kItems], pz[kItems];
#pragma unroll
  for (int k = 0; k < kItems; ++k) {
    const bool valid = full || wbase + k * 32 < n;
    px[k] = valid ? src.cx(wbase + k * 32) : 0.0;
    py[k] = valid ? src.cy(wbase + k * 32) : 0.0;
    pz[k] = valid ? src.cz(wbase + k * 32) : 0.0;
  }
  // the packed keys; the bucket digit = the leading bits of the voxel x index (the same expression tile_digits_kernel counted).
  // The key goes out in input order at once.
  KeyT key[kItems];
  int bad = 0;
#pragma unroll
  for (int k = 0; k < kItems; ++k) {
    int b = 0;
    unsigned long long rx;
    const bool valid = full || wbase + k * 32 < n;
    key[k] = pack_one<KeyT>(px[k], py[k], pz[k], g, b, &rx);
    if (!kDigitBytes) dig[k] = valid ? (unsigned)(rx >> dshift) & mask : (unsigned)(kRadix - 1);
    // out of the grid (status bit) or past the end of the input: keep the key consistent with the digit it is ranked by
    if (b || !valid) key[k] = (KeyT)dig[k] << shift;
    bad |= valid && b;
    if (valid) keys_orig[wbase + k * 32] = key[k];
  }
  __syncthreads();  // rank_init

  unsigned rank2[kItems / 2];
  warp_rank<kItems>([&](int k) { return dig[k]; }, sm.r, sm.u.masks, rank2);
  __syncthreads();

  // first slot of every digit inside the tile, and of every warp's keys of it
  {
    unsigned cnt[kWarps], run = 0;
#pragma unroll
    for (int w = 0; w < kWarps; ++w) {
      cnt[w] = sm.r.warp_cnt[w][tid];
      run += cnt[w];
    }
    unsigned total;
    const unsigned b = block_exclusive_scan(run, sm.r.warp_s, &total);
    unsigned acc = b;
#pragma unroll
    for (int w = 0; w < kWarps; ++w) {
      sm.r.warp_cnt[w][tid] = acc;
      acc += cnt[w];
    }
    sm.r.delta[tid] = my_offset - b;
    __syncthreads();
  }

  // reorder the records inside shared memory (the mask words are dead: every warp passed the barriers above), so
  // that a bucket's run leaves as whole 128-byte lines: scattered single-sector stores were measured at a quarter of
  // the HBM write bandwidth (profiles/).
  const unsigned* wc = sm.r.warp_cnt[warp];
#pragma unroll
  for (int k = 0; k < kItems; ++k) {
    const unsigned pos = wc[dig[k]] + rank_of(rank2, k);
    sm.u.rec.x[pos] = px[k];
    sm.u.rec.y[pos] = py[k];
    sm.u.rec.z[pos] = pz[k];
    sm.u.rec.key[pos] = key[k];
    sm.u.rec.slot[pos] = (unsigned short)(slot0 + k * 32);
  }
  __syncthreads();

  // 32-bit keys also ride in the upper half of the record's index slot (readers only look at the lower half)
  const int tile_n = (int)min((int64_t)kTile, n - tile_base);
#pragma unroll 3
  for (int p = tid; p < tile_n; p += kThreads) {
    const KeyT k = sm.u.rec.key[p];
    unsigned long long w = (uint32_t)tile_base + (uint32_t)sm.u.rec.slot[p];
    if (sizeof(KeyT) == 4) w |= (unsigned long long)k << 32;
    const unsigned dst = (unsigned)p + sm.r.delta[bucket_of(k, shift, mask)];
    st_record(rec_b + dst, sm.u.rec.x[p], sm.u.rec.y[p], sm.u.rec.z[p], __longlong_as_double((long long)w));
    keys_b[dst] = sizeof(KeyOutT) < sizeof(KeyT) ? (KeyOutT)(k & (((KeyT)1 << shift) - 1)) : (KeyOutT)k;
  }
  if (__any_sync(0xffffffffu, bad) && lane == 0) atomicOr(status, 1);
}

// full sorted keys from the sorted low parts: key = bucket << low_bits | low, the bucket of a position from bucket_start
__global__ void __launch_bounds__(kThreads) expand_keys_kernel(const uint32_t* __restrict__ low, int64_t n, const uint32_t* __restrict__ bucket_start,
                                                               int nb, int low_bits, unsigned long long* __restrict__ out) {
  __shared__ unsigned start[kRadix + 1];
  for (int i = threadIdx.x; i <= nb; i += blockDim.x) start[i] = bucket_start[i];
  __syncthreads();
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; p < n; p += stride) {
    int lo = 0, hi = nb;  // start[lo] <= p < start[hi]
    while (hi - lo > 1) {
      const int mid = (lo + hi) >> 1;
      if (start[mid] <= (unsigned)p) lo = mid; else hi = mid;
    }
    out[p] = ((unsigned long long)lo << low_bits) | low[p];
  }
}

inline int64_t num_tiles(int64_t n) { return (n + kTile - 1) / kTile; }
inline size_t align256(size_t v) { return (v + 255) / 256 * 256; }
// workspace: [digits n x 1, only with VAPC_BUCKET_DIGIT_BYTES][tile_counts tiles x 256 x 2][tile_offsets tiles x 256 x 4][chunk sums (chunks + 4 groups) x 256 x 4]
struct BucketWs {
  uint8_t* digits;
  uint16_t* tile_counts;
  uint32_t* tile_offsets;
  uint32_t* chunk_sums;
};
inline int64_t num_chunks(int64_t ntiles) { return (ntiles + kChunkTiles - 1) / kChunkTiles; }
inline size_t bucket_ws_bytes(int64_t n) {
  const int64_t nt = num_tiles(n < 1 ? 1 : n);
  return align256(kDigitBytes ? (size_t)n : 0) + align256((size_t)nt * kRadix * 2) + align256((size_t)nt * kRadix * 4) + align256(((size_t)num_chunks(nt) + kScanGroups) * kRadix * 4);
}
inline BucketWs carve_bucket_ws(void* ws, int64_t n) {
  const int64_t nt = num_tiles(n);
  unsigned char* p = static_cast<unsigned char*>(ws);
  BucketWs w;
  w.digits = p; p += align256(kDigitBytes ? (size_t)n : 0);
  w.tile_counts = reinterpret_cast<uint16_t*>(p); p += align256((size_t)nt * kRadix * 2);
  w.tile_offsets = reinterpret_cast<uint32_t*>(p); p += align256((size_t)nt * kRadix * 4);
  w.chunk_sums = reinterpret_cast<uint32_t*>(p);
  return w;
}

template <typename KeyT, typename KeyOutT, typename Src>
int bucket_impl(const Src& src, int64_t n, const Grid& g, int mb, KeyOutT* keys_b, double4* rec_b, KeyT* keys_orig,
                uint32_t* bucket_start, int32_t* status, void* ws, cudaStream_t s) {
  const int shift = g.bx + g.by + g.bz - mb;  // the bucket digit = the leading mb bits of the key = of the voxel x index
  const unsigned mask = (1u << mb) - 1u;
  const int64_t ntiles = num_tiles(n);
  const int nchunks = (int)num_chunks(ntiles);
  const BucketWs w = carve_bucket_ws(ws, n);
  VAPC_LAUNCH((tile_digits_kernel<Src>), (unsigned)ntiles, kThreads, 0, s, src, n, g, g.bx - mb, mask, w.digits, w.tile_counts, status);
  VAPC_LAUNCH(tile_chunk_sums_kernel, (unsigned)nchunks, kThreads, 0, s, w.tile_counts, ntiles, w.chunk_sums);
  const int per = (nchunks + kScanGroups - 1) / kScanGroups;  // chunks per scan group
  uint32_t* group_add = w.chunk_sums + (size_t)nchunks * kRadix;
  VAPC_LAUNCH(tile_chunk_scan_kernel, 1, kThreads * kScanGroups, 0, s, w.chunk_sums, nchunks, per, 1 << mb, (uint32_t)n, group_add, bucket_start);
  VAPC_LAUNCH(tile_offsets_kernel, (unsigned)nchunks, kThreads, 0, s, w.tile_counts, w.chunk_sums, group_add, per, ntiles, w.tile_offsets);
  const size_t smem = sizeof(BucketSmem<KeyT>);
  VAPC_RETURN_IF_CUDA(cudaFuncSetAttribute(tile_scatter_kernel<KeyT, KeyOutT, Src>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  VAPC_LAUNCH((tile_scatter_kernel<KeyT, KeyOutT, Src>), (unsigned)ntiles, kThreads, smem, s, w.digits, src, n, g, shift, g.bx - mb, mask, w.tile_offsets, keys_orig, keys_b,
              rec_b, status);
  VAPC_CHECK_LAUNCH();
  return VAPC_OK;
}
}  // namespace

extern "C" size_t vapc_bucket_workspace_bytes(int64_t n) { return bucket_ws_bytes(n < 1 ? 1 : n); }

namespace {
template <typename Src>
int pack_bucketed(const char* who, const Src& src, int64_t n, const vapc_grid_t* grid_host, void* keys_bucketed, double* xyzw_bucketed, void* keys_orig,
                  uint32_t* bucket_start, int32_t* bucket_bits_host, int32_t* bucketed_key_bytes_host, int32_t* status, void* ws, size_t ws_bytes,
                  void* stream) {
  if (!grid_host || !(grid_host->voxel_size > 0)) return vapc_set_error(VAPC_E_INVALID, "%s: bad grid", who);
  if (n < 0 || !bucket_bits_host || !bucketed_key_bytes_host) return vapc_set_error(VAPC_E_INVALID, "%s: bad arguments", who);
  const Grid g = to_device_grid(grid_host);
  const int mb = g.bx < kRadixBits ? g.bx : kRadixBits;
  const int low_bits = g.bx + g.by + g.bz - mb;
  const bool narrow = grid_host->key_bytes == 8 && low_bits <= 32;  // the bits below the bucket digit fit a 32-bit word
  *bucket_bits_host = mb;
  *bucketed_key_bytes_host = narrow ? 4 : grid_host->key_bytes;
  if (n == 0) return VAPC_OK;
  if (n >= (int64_t)1 << 32) return vapc_set_error(VAPC_E_INVALID, "%s: n must be < 2^32 per GPU", who);
  if (src.null() || !keys_bucketed || !xyzw_bucketed || !keys_orig || !bucket_start || !status) return vapc_set_error(VAPC_E_INVALID, "%s: null buffer", who);
  if (reinterpret_cast<uintptr_t>(xyzw_bucketed) & 31u) return vapc_set_error(VAPC_E_INVALID, "%s: xyzw must be 32-byte aligned", who);
  if (!ws || ws_bytes < vapc_bucket_workspace_bytes(n)) return vapc_set_error(VAPC_E_WORKSPACE, "%s: workspace too small", who);
  cudaStream_t s = as_stream(stream);
  double4* rec = reinterpret_cast<double4*>(xyzw_bucketed);
  typedef unsigned long long u64;
  typedef uint32_t u32;
  if (grid_host->key_bytes == 4) {
    return bucket_impl<u32, u32>(src, n, g, mb, static_cast<u32*>(keys_bucketed), rec, static_cast<u32*>(keys_orig), bucket_start, status, ws, s);
  }
  if (grid_host->key_bytes == 8) {
    if (narrow) {
      return bucket_impl<u64, u32>(src, n, g, mb, static_cast<u32*>(keys_bucketed), rec, static_cast<u64*>(keys_orig), bucket_start, status, ws, s);
    }
    return bucket_impl<u64, u64>(src, n, g, mb, static_cast<u64*>(keys_bucketed), rec, static_cast<u64*>(keys_orig), bucket_start, status, ws, s);
  }
  return vapc_set_error(VAPC_E_INVALID, "%s: key_bytes must be 4 or 8", who);
}
}  // namespace

extern "C" int vapc_pack_records_bucketed(const double* x, const double* y, const double* z, int64_t n, const vapc_grid_t* grid_host,
                                          void* keys_bucketed, double* xyzw_bucketed, void* keys_orig, uint32_t* bucket_start,
                                          int32_t* bucket_bits_host, int32_t* bucketed_key_bytes_host, int32_t* status, void* ws, size_t ws_bytes,
                                          void* stream) {
  return pack_bucketed("vapc_pack_records_bucketed", SrcF64{x, y, z}, n, grid_host, keys_bucketed, xyzw_bucketed, keys_orig, bucket_start, bucket_bits_host,
                       bucketed_key_bytes_host, status, ws, ws_bytes, stream);
}

extern "C" int vapc_pack_records_bucketed_i32(const int32_t* X, const int32_t* Y, const int32_t* Z, int64_t n, const vapc_las_transform_t* t_host,
                                              const vapc_grid_t* grid_host, void* keys_bucketed, double* xyzw_bucketed, void* keys_orig,
                                              uint32_t* bucket_start, int32_t* bucket_bits_host, int32_t* bucketed_key_bytes_host, int32_t* status,
                                              void* ws, size_t ws_bytes, void* stream) {
  if (!t_host) return vapc_set_error(VAPC_E_INVALID, "vapc_pack_records_bucketed_i32: null transform");
  return pack_bucketed("vapc_pack_records_bucketed_i32", make_src_i32(X, Y, Z, t_host), n, grid_host, keys_bucketed, xyzw_bucketed, keys_orig, bucket_start,
                       bucket_bits_host, bucketed_key_bytes_host, status, ws, ws_bytes, stream);
}

extern "C" int vapc_expand_bucket_keys(const uint32_t* low_sorted, int64_t n, const uint32_t* bucket_start, int32_t bucket_bits, int32_t low_bits,
                                       uint64_t* keys_out, void* stream) {
  if (n < 0 || bucket_bits < 0 || bucket_bits > kRadixBits || low_bits < 0 || low_bits > 32)
    return vapc_set_error(VAPC_E_INVALID, "vapc_expand_bucket_keys: bad arguments");
  if (n == 0) return VAPC_OK;
  if (!low_sorted || !bucket_start || !keys_out) return vapc_set_error(VAPC_E_INVALID, "vapc_expand_bucket_keys: null buffer");
  const int grid = (int)std::max<int64_t>(1, std::min<int64_t>((n + kThreads * 4 - 1) / (kThreads * 4), (int64_t)kNumSMs * 16));
  VAPC_LAUNCH(expand_keys_kernel, grid, kThreads, 0, as_stream(stream), low_sorted, n, bucket_start, 1 << bucket_bits, low_bits,
              reinterpret_cast<unsigned long long*>(keys_out));
  VAPC_CHECK_LAUNCH();
  return VAPC_OK;
}
