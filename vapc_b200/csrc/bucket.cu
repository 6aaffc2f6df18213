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
//   keys_hist_kernel      reads x, y, z; writes the packed keys in input order and counts the bucket digit
//                         (shared-memory atomics)                                              24 + K B/pt
//   bucket_scatter_kernel reads the keys of a 4096-point tile, ranks the bucket digit (radix_rank.cuh, stable),
//                         decoupled look-back for the global offsets, then reads x, y, z and writes key and
//                         record {x, y, z, index} at the bucketed position, records straight from registers as one
//                         256-bit store each (a full sector)                              K + 24 + K + 32 B/pt
// (One fused kernel was tried first: with the three IEEE divisions of the key in front of the count publication
// every tile waited on its predecessors' look-back words — 34 % of all stall samples on that poll, 4.4 ms per 1e8
// points.  Keys first, then a scatter pass whose tiles publish right after a 16 KB key load, is 2x faster.)
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
constexpr int kItems = VAPC_BUCKET_ITEMS;
constexpr int kTile = kThreads * kItems;  // 1536 points: the reordered records of a tile take 48 KB of shared memory (6 vs 8 items: 2.22 vs 2.42 ms)

template <typename KeyT>
__device__ __forceinline__ unsigned bucket_of(KeyT key, int shift, unsigned mask) { return (unsigned)(key >> shift) & mask; }

// keys in input order + histogram of the bucket digit
template <typename KeyT, int VEC, typename Src>
__global__ void __launch_bounds__(kThreads) keys_hist_kernel(Src src, int64_t n, Grid g, int shift, unsigned mask, KeyT* __restrict__ keys,
                                                             uint32_t* __restrict__ hist, int* __restrict__ status) {
  __shared__ unsigned sh[kRadix];
  sh[threadIdx.x] = 0;
  __syncthreads();
  int bad = 0;
  const int64_t nvec = n / VEC;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t v = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; v < nvec; v += stride) {
    if (VEC == 2) {
      double a[2], b[2], c[2];
      src.pair(v, a, b, c);
      const KeyT k0 = pack_one<KeyT>(a[0], b[0], c[0], g, bad);
      const KeyT k1 = pack_one<KeyT>(a[1], b[1], c[1], g, bad);
      if (sizeof(KeyT) == 4) {
        reinterpret_cast<uint2*>(keys)[v] = make_uint2((unsigned)k0, (unsigned)k1);
      } else {
        reinterpret_cast<ulonglong2*>(keys)[v] = make_ulonglong2(k0, k1);
      }
      atomicAdd(&sh[bucket_of(k0, shift, mask)], 1u);
      atomicAdd(&sh[bucket_of(k1, shift, mask)], 1u);
    } else {
      const KeyT k = pack_one<KeyT>(src.cx(v), src.cy(v), src.cz(v), g, bad);
      keys[v] = k;
      atomicAdd(&sh[bucket_of(k, shift, mask)], 1u);
    }
  }
  if (VEC == 2 && (n & 1) && blockIdx.x == 0 && threadIdx.x == 0) {
    const KeyT k = pack_one<KeyT>(src.cx(n - 1), src.cy(n - 1), src.cz(n - 1), g, bad);
    keys[n - 1] = k;
    atomicAdd(&sh[bucket_of(k, shift, mask)], 1u);
  }
  __syncthreads();
  const unsigned c = sh[threadIdx.x];
  if (c) atomicAdd(&hist[threadIdx.x], c);
  if (__any_sync(0xffffffffu, bad) && (threadIdx.x & 31) == 0) atomicOr(status, 1);
}

// bucket_start[b] = first slot of bucket b, [nb] = n   (digit_base = exclusive scan of the histogram)
__global__ void bucket_start_kernel(const uint32_t* __restrict__ digit_base, int nb, uint32_t n, uint32_t* __restrict__ bucket_start) {
  const int t = threadIdx.x;
  if (t < nb) bucket_start[t] = digit_base[t];
  if (t == 0) bucket_start[nb] = n;
}

// The tile's records reordered by bucket, as columns: coordinates, key and the 16-bit input slot of every point (64-bit
// shared-memory stores per column measured 4 % faster than one 32-byte record store; a fourth CTA per SM — 55 KB each —
// needs 64 registers per thread and loses more to spills than it gains).
#ifndef VAPC_BUCKET_TMA
#define VAPC_BUCKET_TMA 0
#endif
template <typename KeyT>
struct BucketSmem {
  RankSmem r;
  alignas(32) union {
    MaskBuffers masks;
#if VAPC_BUCKET_TMA
    struct {
      double4 rec[kTile];  // whole records in bucket order: a bucket's run leaves as ONE bulk copy (TMA) instead of 32-byte stores
      KeyT key[kTile];
    } rec;
#else
    struct {
      double x[kTile], y[kTile], z[kTile];
      KeyT key[kTile];
      unsigned short slot[kTile];
    } rec;
#endif
  } u;
};

// KeyOutT: the bucketed keys are written as 32-bit words holding the key bits BELOW the bucket digit when those fit
// (64-bit keys of up to 40 bits): the segmented sort then moves 4-byte keys; vapc_expand_bucket_keys restores the full keys.
template <typename KeyT, typename KeyOutT, typename LookT, typename Src>
__global__ void __launch_bounds__(kThreads, sizeof(KeyT) == 4 ? VAPC_BUCKET_MINB : 3) bucket_scatter_kernel(
    const KeyT* __restrict__ keys, Src src, int64_t n, int shift,
    unsigned mask, const uint32_t* __restrict__ digit_base, LookT* __restrict__ look, unsigned* __restrict__ ticket, KeyOutT* __restrict__ keys_b,
    double4* __restrict__ rec_b) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  BucketSmem<KeyT>& sm = *reinterpret_cast<BucketSmem<KeyT>*>(smem_raw);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (tid == 0) sm.r.tile = atomicAdd(ticket, 1u);
  rank_init(sm.r, sm.u.masks);
  __syncthreads();
  const unsigned tile = sm.r.tile;
  const int64_t tile_base = (int64_t)tile * kTile;
  const bool full = tile_base + kTile <= n;

  // warp-striped: warp w owns points [w*32*kItems, (w+1)*32*kItems) of the tile, lane l takes point k*32 + l in step k.
  // Slots past the end of the input (last tile) hold all-ones keys: highest bucket, last ranks, never written.
  const int slot0 = warp * (32 * kItems) + lane;
  const int64_t wbase = tile_base + slot0;
  KeyT key[kItems];
#pragma unroll
  for (int k = 0; k < kItems; ++k) key[k] = (full || wbase + k * 32 < n) ? __ldcs(keys + wbase + k * 32) : (KeyT)~(KeyT)0;
  // the coordinates are only needed once the positions are known; their loads overlap the ranking
  double px[kItems], py[kItems], pz[kItems];
#pragma unroll
  for (int k = 0; k < kItems; ++k) {
    const bool valid = full || wbase + k * 32 < n;
    px[k] = valid ? src.cx(wbase + k * 32) : 0.0;
    py[k] = valid ? src.cy(wbase + k * 32) : 0.0;
    pz[k] = valid ? src.cz(wbase + k * 32) : 0.0;
  }

  unsigned rank2[kItems / 2];
  warp_rank<kItems>([&](int k) { return bucket_of(key[k], shift, mask); }, sm.r, sm.u.masks, rank2);
  __syncthreads();

  LookT* mine = look + (int64_t)tile * kRadix + tid;
  unsigned bstart;
  const unsigned run = digit_offsets<LookT>(sm.r, mine, tile == 0, &bstart);

  // reorder the records inside shared memory (the mask words are dead: every warp passed the barriers above), so
  // that a bucket's run leaves as whole 128-byte lines: scattered single-sector stores were measured at a quarter of
  // the HBM write bandwidth (profiles/).
  const unsigned* wc = sm.r.warp_cnt[warp];
#if VAPC_BUCKET_TMA
#pragma unroll
  for (int k = 0; k < kItems; ++k) {
    const unsigned pos = wc[bucket_of(key[k], shift, mask)] + rank_of(rank2, k);
    unsigned long long w = (uint32_t)tile_base + (uint32_t)(slot0 + k * 32);
    if (sizeof(KeyT) == 4) w |= (unsigned long long)key[k] << 32;
    sm.u.rec.rec[pos] = make_double4(px[k], py[k], pz[k], __longlong_as_double((long long)w));
    sm.u.rec.key[pos] = key[k];
  }
  const LookT excl = look_back<LookT>(mine, (int64_t)tile, run);
  sm.r.delta[tid] = digit_base[tid] + (unsigned)excl - bstart;
  // the records were written through the generic proxy; the bulk copies read them through the async proxy
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  __syncthreads();
  const int tile_n = (int)min((int64_t)kTile, n - tile_base);
  {
    // thread == bucket: its run [bstart, bstart + run) of the tile goes to rec_b[digit_base + excl ...] as one bulk copy
    const int end = min((int)(bstart + run), tile_n);
    if (end > (int)bstart) {
      const unsigned bytes = (unsigned)(end - (int)bstart) * 32u;
      const unsigned src = (unsigned)__cvta_generic_to_shared(&sm.u.rec.rec[bstart]);
      double4* dst = rec_b + (digit_base[tid] + (unsigned)excl);
      asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(src), "r"(bytes) : "memory");
    }
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
  }
#pragma unroll 3
  for (int p = tid; p < tile_n; p += kThreads) {
    const KeyT k = sm.u.rec.key[p];
    const unsigned dst = (unsigned)p + sm.r.delta[bucket_of(k, shift, mask)];
    keys_b[dst] = sizeof(KeyOutT) < sizeof(KeyT) ? (KeyOutT)(k & (((KeyT)1 << shift) - 1)) : (KeyOutT)k;
  }
  asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");  // shared memory must outlive the copies
#else
#pragma unroll
  for (int k = 0; k < kItems; ++k) {
    const unsigned pos = wc[bucket_of(key[k], shift, mask)] + rank_of(rank2, k);
    sm.u.rec.x[pos] = px[k];
    sm.u.rec.y[pos] = py[k];
    sm.u.rec.z[pos] = pz[k];
    sm.u.rec.key[pos] = key[k];
    sm.u.rec.slot[pos] = (unsigned short)(slot0 + k * 32);
  }

  const LookT excl = look_back<LookT>(mine, (int64_t)tile, run);
  sm.r.delta[tid] = digit_base[tid] + (unsigned)excl - bstart;
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
#endif
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
// workspace: [bucket histogram -> first slots, 256 u32][ticket (256 B)][look-back ntiles x 256 x 8 B]
constexpr size_t kHistBytes = kRadix * sizeof(uint32_t);

template <typename KeyT, typename KeyOutT, typename LookT, typename Src>
int bucket_impl(const Src& src, int64_t n, const Grid& g, int mb, KeyOutT* keys_b, double4* rec_b, KeyT* keys_orig,
                uint32_t* bucket_start, int32_t* status, void* ws, cudaStream_t s) {
  const int shift = g.bx + g.by + g.bz - mb;  // the bucket digit = the leading mb bits of the key
  const unsigned mask = (1u << mb) - 1u;
  const int64_t ntiles = num_tiles(n);
  unsigned char* base = static_cast<unsigned char*>(ws);
  uint32_t* hist = reinterpret_cast<uint32_t*>(base);
  unsigned* ticket = reinterpret_cast<unsigned*>(base + kHistBytes);
  LookT* look = reinterpret_cast<LookT*>(base + kHistBytes + 256);
  VAPC_RETURN_IF_CUDA(cudaMemsetAsync(base, 0, kHistBytes + 256 + (size_t)ntiles * kRadix * sizeof(LookT), s));
  const int hgrid = (int)std::max<int64_t>(1, std::min<int64_t>((n + kThreads * 4 - 1) / (kThreads * 4), (int64_t)kNumSMs * 16));
  const bool vec = src.aligned() && (reinterpret_cast<uintptr_t>(keys_orig) & 15u) == 0;
  if (vec) VAPC_LAUNCH((keys_hist_kernel<KeyT, 2, Src>), hgrid, kThreads, 0, s, src, n, g, shift, mask, keys_orig, hist, status);
  else VAPC_LAUNCH((keys_hist_kernel<KeyT, 1, Src>), hgrid, kThreads, 0, s, src, n, g, shift, mask, keys_orig, hist, status);
  VAPC_LAUNCH(digit_scan_kernel, 1, kThreads, 0, s, hist, nullptr, 1);
  VAPC_LAUNCH(bucket_start_kernel, 1, kThreads, 0, s, hist, 1 << mb, (uint32_t)n, bucket_start);
  const size_t smem = sizeof(BucketSmem<KeyT>);
  VAPC_RETURN_IF_CUDA(cudaFuncSetAttribute(bucket_scatter_kernel<KeyT, KeyOutT, LookT, Src>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  VAPC_LAUNCH((bucket_scatter_kernel<KeyT, KeyOutT, LookT, Src>), (unsigned)ntiles, kThreads, smem, s, keys_orig, src, n, shift, mask, hist, look, ticket, keys_b,
              rec_b);
  VAPC_CHECK_LAUNCH();
  return VAPC_OK;
}
}  // namespace

extern "C" size_t vapc_bucket_workspace_bytes(int64_t n) {
  if (n < 1) n = 1;
  return kHistBytes + 256 + (size_t)num_tiles(n) * kRadix * sizeof(unsigned long long);
}

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
  const bool wide = n >= ((int64_t)1 << 30);
  typedef unsigned long long u64;
  typedef uint32_t u32;
  if (grid_host->key_bytes == 4) {
    if (!wide) return bucket_impl<u32, u32, u32>(src, n, g, mb, static_cast<u32*>(keys_bucketed), rec, static_cast<u32*>(keys_orig), bucket_start, status, ws, s);
    return bucket_impl<u32, u32, u64>(src, n, g, mb, static_cast<u32*>(keys_bucketed), rec, static_cast<u32*>(keys_orig), bucket_start, status, ws, s);
  }
  if (grid_host->key_bytes == 8) {
    if (narrow) {
      if (!wide) return bucket_impl<u64, u32, u32>(src, n, g, mb, static_cast<u32*>(keys_bucketed), rec, static_cast<u64*>(keys_orig), bucket_start, status, ws, s);
      return bucket_impl<u64, u32, u64>(src, n, g, mb, static_cast<u32*>(keys_bucketed), rec, static_cast<u64*>(keys_orig), bucket_start, status, ws, s);
    }
    if (!wide) return bucket_impl<u64, u64, u32>(src, n, g, mb, static_cast<u64*>(keys_bucketed), rec, static_cast<u64*>(keys_orig), bucket_start, status, ws, s);
    return bucket_impl<u64, u64, u64>(src, n, g, mb, static_cast<u64*>(keys_bucketed), rec, static_cast<u64*>(keys_orig), bucket_start, status, ws, s);
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
