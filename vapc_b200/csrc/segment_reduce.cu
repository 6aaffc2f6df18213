// Run-length segmentation of the sorted keys + single-pass segmented reductions.
//
// A tile is kSegTile = 1024 consecutive sorted positions handled by one CTA of 256 threads
// (4 positions per thread, blocked, so key / index loads are 16-byte vectors).  In a tile:
//   1. head flags (key differs from its predecessor), block scan -> local segment ids;
//   2. every position gathers its 32-byte xyzw record (one sector per point, all loads issued
//      up front) into shared memory; point->voxel ids are scattered out;
//   3. one thread per segment walks its points in shared memory IN SORTED ORDER — which is
//      original point order, the sort being stable — accumulating n, SUM d, SUM d d' with
//      d = (p - c) - (p_first - c), c = c(voxel), i.e. shifted by the segment's first point so that
//      the final M2 = SUM dd' - SUM d SUM d'/n loses nothing to cancellation even for a tight
//      cluster in a large voxel; the segment leaves as (n, mean of p - c, M2 about the mean).
//      Segments that start in the tile are stored to the voxel row, the leading partial segment
//      (a voxel continued from the previous tile) goes to a per-tile carry slot;
//   4. a fix-up kernel folds each voxel's carry chain in with Chan's pairwise update, in tile order.
// (n, mean - c, M2) triples combine exactly the same way across GPUs (vapc_merge_partials).
// No atomics: results are bit-reproducible run to run.
//
// Algorithmic bytes per point (u32 keys): 4 key + 4 idx + 32 record + 4 point2voxel write(+28 of
// sector write-allocate on the scatter) ; per voxel: 8 key... see DESIGN.md.
#include <math.h>

#include <algorithm>

#include "common.cuh"
#include "math3.cuh"
#include "segment_ws.cuh"
#include "stats_row.cuh"

namespace {
using namespace segws;
constexpr int kThreads = 256;
constexpr int kItems = 4;
static_assert(kSegTile == kThreads * kItems, "tile = 4 positions per thread");
constexpr int kPartialRowWords = 11;         // exchanged partial voxel row: key, count, mean[3], M2[6] (8 bytes each)
#ifndef VAPC_REDUCE_MINB
#define VAPC_REDUCE_MINB 4
#endif

// ---- load one tile's keys (blocked, 4 per thread) + the predecessor of the first ---------------
template <typename KeyT>
__device__ __forceinline__ void load_tile_keys(const KeyT* __restrict__ keys, int64_t n, int64_t base, KeyT (&key)[kItems],
                                               KeyT& prev, bool (&valid)[kItems]) {
  const int64_t p0 = base + (int64_t)threadIdx.x * kItems;
  if (p0 + kItems <= n) {
    if (sizeof(KeyT) == 4) {
      const uint4 v = *reinterpret_cast<const uint4*>(keys + p0);
      key[0] = (KeyT)v.x; key[1] = (KeyT)v.y; key[2] = (KeyT)v.z; key[3] = (KeyT)v.w;
    } else {
      const ulonglong2 a = *reinterpret_cast<const ulonglong2*>(keys + p0);
      const ulonglong2 b = *reinterpret_cast<const ulonglong2*>(keys + p0 + 2);
      key[0] = (KeyT)a.x; key[1] = (KeyT)a.y; key[2] = (KeyT)b.x; key[3] = (KeyT)b.y;
    }
#pragma unroll
    for (int j = 0; j < kItems; ++j) valid[j] = true;
  } else {
#pragma unroll
    for (int j = 0; j < kItems; ++j) {
      valid[j] = p0 + j < n;
      key[j] = valid[j] ? keys[p0 + j] : (KeyT)0;
    }
  }
  // predecessor of this thread's first position: previous thread's last key (via shuffle /
  // shared memory would need a barrier; a global re-read of one L1-resident word is simpler)
  prev = (p0 > 0 && p0 <= n) ? keys[p0 - 1] : (KeyT)0;
}

template <typename KeyT>
__device__ __forceinline__ unsigned head_flags(const KeyT (&key)[kItems], KeyT prev, const bool (&valid)[kItems], int64_t p0,
                                               bool (&head)[kItems]) {
  unsigned cnt = 0;
#pragma unroll
  for (int j = 0; j < kItems; ++j) {
    const KeyT before = j == 0 ? prev : key[j - 1];
    head[j] = valid[j] && ((p0 + j) == 0 || key[j] != before);
    cnt += head[j] ? 1u : 0u;
  }
  return cnt;
}

// ---- step 1: heads per tile ---------------------------------------------------------------------
template <typename KeyT>
__global__ void __launch_bounds__(kThreads) count_heads_kernel(const KeyT* __restrict__ keys, int64_t n,
                                                               uint32_t* __restrict__ tile_heads) {
  __shared__ unsigned warp_cnt[kThreads / 32];
  const int64_t base = (int64_t)blockIdx.x * kSegTile;
  KeyT key[kItems], prev;
  bool valid[kItems], head[kItems];
  load_tile_keys(keys, n, base, key, prev, valid);
  unsigned c = head_flags(key, prev, valid, base + (int64_t)threadIdx.x * kItems, head);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
  if ((threadIdx.x & 31) == 0) warp_cnt[threadIdx.x >> 5] = c;
  __syncthreads();
  if (threadIdx.x == 0) {
    unsigned t = 0;
    for (int w = 0; w < kThreads / 32; ++w) t += warp_cnt[w];
    tile_heads[blockIdx.x] = t;
  }
}

// ---- shared tile state ----------------------------------------------------------------------------
#ifndef VAPC_LONG_SEG
#define VAPC_LONG_SEG 64
#endif
__device__ __forceinline__ void prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }
#ifndef VAPC_MID_SEG
#define VAPC_MID_SEG 16
#endif
#ifndef VAPC_MID_MEAN
#define VAPC_MID_MEAN 8
#endif
constexpr int kLongSeg = VAPC_LONG_SEG;  // local segments of this many points or more are reduced by a warp instead of one thread
constexpr int kMidSeg = VAPC_MID_SEG;    // ... and, in the kernel for clouds of fewer than VAPC_MID_MEAN points per voxel, those of kMidSeg ..
                                         // kLongSeg - 1 points by a group of 8 lanes
constexpr int kQueueSeg = kMidSeg;
template <typename KeyT>
struct TileSmem {
  unsigned scan_s[33];
  unsigned short seg_start[kSegTile + 2];  // [s] first tile position of local segment s; [nheads+1] = tile_n
  KeyT seg_key[kSegTile + 1];              // key of local segment s (s = 0: the continued voxel)
  // coordinates of tile position p live at slot sw(p): threads store 4 consecutive positions each, i.e. lane l writes
  // position 4 l + j — unpadded, a half-warp's 64-bit stores hit only 4 of the 16 bank pairs (4-way conflict, 8 wavefronts
  // per store instead of 2: 5.6e7 of the kernel's 1.6e8 shared-memory wavefronts at 1e8 points, profiles/r1_d).  One pad
  // slot per 16 positions makes 4 l + j + l / 4 distinct mod 16.
  double x[kSegTile + kSegTile / 16], y[kSegTile + kSegTile / 16], z[kSegTile + kSegTile / 16];
  // local segments of >= kLongSeg points are not walked by one thread but handed to a warp each
  unsigned n_long;
  unsigned short long_seg[kSegTile / kLongSeg + 2];
  unsigned n_mid;
  unsigned short mid_seg[kSegTile / kQueueSeg + 2];
};
__device__ __forceinline__ int sw(int p) { return p + (p >> 4); }

template <typename KeyT>
__device__ __forceinline__ void tile_prologue_loaded(TileSmem<KeyT>& sm, int64_t n, int64_t base, KeyT (&key)[kItems], KeyT prev,
                                                     bool (&valid)[kItems], bool (&head)[kItems], unsigned (&inc)[kItems], unsigned& nheads,
                                                     int& tile_n);
// Common prologue: flags, scan, segment tables, voxel-table bookkeeping.
// On return inc[j] = number of heads at tile positions <= this one (local segment id; 0 = continuation).
template <typename KeyT>
__device__ __forceinline__ void tile_prologue(TileSmem<KeyT>& sm, const KeyT* __restrict__ keys, int64_t n, int64_t base,
                                              KeyT (&key)[kItems], bool (&valid)[kItems], bool (&head)[kItems],
                                              unsigned (&inc)[kItems], unsigned& nheads, int& tile_n) {
  KeyT prev;
  load_tile_keys(keys, n, base, key, prev, valid);
  tile_prologue_loaded(sm, n, base, key, prev, valid, head, inc, nheads, tile_n);
}
// The same for keys the caller has loaded already (load_tile_keys): a kernel that also gathers through a second array issues both
// loads first and runs this scan while the gathers are in flight.
template <typename KeyT>
__device__ __forceinline__ void tile_prologue_loaded(TileSmem<KeyT>& sm, int64_t n, int64_t base, KeyT (&key)[kItems], KeyT prev,
                                                     bool (&valid)[kItems], bool (&head)[kItems], unsigned (&inc)[kItems], unsigned& nheads,
                                                     int& tile_n) {
  const int64_t p0 = base + (int64_t)threadIdx.x * kItems;
  const unsigned local = head_flags(key, prev, valid, p0, head);
  const unsigned excl = block_exclusive_scan(local, sm.scan_s, &nheads);
  unsigned run = excl;
  tile_n = (int)min((int64_t)kSegTile, n - base);
#pragma unroll
  for (int j = 0; j < kItems; ++j) {
    run += head[j] ? 1u : 0u;
    inc[j] = run;
    if (head[j]) {
      sm.seg_start[run] = (unsigned short)(threadIdx.x * kItems + j);
      sm.seg_key[run] = key[j];
    }
  }
  if (threadIdx.x == 0) {
    sm.seg_start[0] = 0;
    sm.seg_key[0] = key[0];  // position 0 belongs to segment 0 (continuation) or is head of segment 1
    sm.seg_start[nheads + 1] = (unsigned short)tile_n;
  }
}

// ---- weighted run lengths of sorted keys (multi-GPU mode / mode_count: (voxel, value) runs instead of points) -------
// run_keys[r] = the r-th distinct key, run_weight[r] += the weights of its elements.  64-bit integer atomics: exact and
// independent of the order of arrival.  tile_first = the workspace vapc_count_voxels left behind.
template <typename KeyT>
__global__ void __launch_bounds__(kThreads) run_lengths_kernel(const KeyT* __restrict__ keys, const uint32_t* __restrict__ weights, int64_t n,
                                                               const uint32_t* __restrict__ tile_first, KeyT* __restrict__ run_keys,
                                                               unsigned long long* __restrict__ run_weight) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  TileSmem<KeyT>& sm = *reinterpret_cast<TileSmem<KeyT>*>(smem_raw);
  const int64_t base = (int64_t)blockIdx.x * kSegTile;
  KeyT key[kItems];
  bool valid[kItems], head[kItems];
  unsigned inc[kItems], nheads;
  int tile_n;
  tile_prologue(sm, keys, n, base, key, valid, head, inc, nheads, tile_n);
  const uint32_t first_row = tile_first[blockIdx.x];
  const int64_t p0 = base + (int64_t)threadIdx.x * kItems;
  unsigned long long acc = 0;
  uint32_t cur = 0xffffffffu;
#pragma unroll
  for (int j = 0; j < kItems; ++j) {
    if (!valid[j]) continue;
    const uint32_t row = first_row + inc[j] - 1u;
    if (head[j]) run_keys[row] = key[j];
    if (row != cur) {
      if (cur != 0xffffffffu) atomicAdd(run_weight + cur, acc);
      cur = row;
      acc = 0;
    }
    acc += weights ? weights[p0 + j] : 1u;
  }
  if (cur != 0xffffffffu) atomicAdd(run_weight + cur, acc);
}

// ---- step 2: moments ------------------------------------------------------------------------------
// MID: the middle tier is compiled in.  Clouds of few points per voxel on average are the ones where single voxels stand out (measured,
// 1.25e8 clustered points, 3.2 per voxel: 2.31 -> 1.97 ms); on a uniform cloud of 10 points per voxel the tier has nothing to do and its
// mere presence cost 0.06 ms of 1.08, so the host picks the kernel from n / V.
template <typename KeyT, bool MID, int L>
__global__ void __launch_bounds__(kThreads, VAPC_REDUCE_MINB) reduce_moments_kernel(
    const KeyT* __restrict__ keys, const uint32_t* __restrict__ pos_sorted, int64_t n, Grid g, const double4* __restrict__ xyzw,
    const uint32_t* __restrict__ tile_first, int64_t nvox, KeyT* __restrict__ voxel_keys, uint32_t* __restrict__ start,
    uint32_t* __restrict__ count, uint32_t* __restrict__ first_idx, double* __restrict__ mean3, double* __restrict__ m2,
    uint32_t* __restrict__ row_sorted, uint32_t* __restrict__ idx_sorted, uint32_t* __restrict__ carry_cnt, double* __restrict__ carry) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  TileSmem<KeyT>& sm = *reinterpret_cast<TileSmem<KeyT>*>(smem_raw);
  const int64_t base = (int64_t)blockIdx.x * kSegTile;
  KeyT key[kItems];
  bool valid[kItems], head[kItems];
  unsigned inc[kItems], nheads;
  int tile_n;

  // the keys, the tile's first row and the record positions are requested together (one memory latency, not two: the scan below
  // then runs while the gathers that depend on the positions are in flight)
  KeyT prev;
  load_tile_keys(keys, n, base, key, prev, valid);
  const uint32_t first_row = tile_first[blockIdx.x];  // voxel row of the first head in this tile
  const int64_t p0 = base + (int64_t)threadIdx.x * kItems;
  uint32_t idx[kItems];
  if (p0 + kItems <= n) {
    const uint4 v = *reinterpret_cast<const uint4*>(pos_sorted + p0);
    idx[0] = v.x; idx[1] = v.y; idx[2] = v.z; idx[3] = v.w;
  } else {
#pragma unroll
    for (int j = 0; j < kItems; ++j) idx[j] = (p0 + j < n) ? pos_sorted[p0 + j] : 0u;
  }
  double4 rec[kItems];
#pragma unroll
  for (int j = 0; j < kItems; ++j) {
    if (p0 + j < n) {
      rec[j] = ld_record(xyzw + idx[j]);
    } else {
      rec[j] = make_double4(0, 0, 0, 0);
    }
  }
  // the record carries the point's original index (bit pattern in w); from here on idx[] is that index
#pragma unroll
  for (int j = 0; j < kItems; ++j) idx[j] = record_index(rec[j]);
  if (idx_sorted) {
    if (p0 + kItems <= n) {
      *reinterpret_cast<uint4*>(idx_sorted + p0) = make_uint4(idx[0], idx[1], idx[2], idx[3]);
    } else {
#pragma unroll
      for (int j = 0; j < kItems; ++j)
        if (p0 + j < n) idx_sorted[p0 + j] = idx[j];
    }
  }

  tile_prologue_loaded(sm, n, base, key, prev, valid, head, inc, nheads, tile_n);

  uint32_t rows[kItems];
#pragma unroll
  for (int j = 0; j < kItems; ++j) {
    const int p = threadIdx.x * kItems + j;
    sm.x[sw(p)] = rec[j].x; sm.y[sw(p)] = rec[j].y; sm.z[sw(p)] = rec[j].z;
    rows[j] = first_row + inc[j] - 1u;  // inc == 0 -> first_row - 1: the continued voxel
    if (valid[j] && head[j]) {
      voxel_keys[rows[j]] = key[j];
      start[rows[j]] = (uint32_t)(p0 + j);
      first_idx[rows[j]] = idx[j];
    }
  }
  if (row_sorted) {  // voxel row of every sorted position (vapc_point2voxel_sorted turns it into the point -> voxel map)
    if (p0 + kItems <= n) {
      *reinterpret_cast<uint4*>(row_sorted + p0) = make_uint4(rows[0], rows[1], rows[2], rows[3]);
    } else {
#pragma unroll
      for (int j = 0; j < kItems; ++j)
        if (valid[j]) row_sorted[p0 + j] = rows[j];
    }
  }
  if (blockIdx.x == gridDim.x - 1 && threadIdx.x == 0) start[nvox] = (uint32_t)n;
  if (threadIdx.x == 0) sm.n_long = sm.n_mid = 0;
  __syncthreads();

  // sums about the segment's first point q0 -> (n, mean of p - c, M2) -> voxel row, or the tile's carry slot for s == 0
  auto emit = [&](unsigned s, int cnt, double q0x, double q0y, double q0z, double sx, double sy, double sz, double xx, double xy, double xz,
                  double yy, double yz, double zz) {
    const double inv = 1.0 / (double)cnt;
    const double mx = sx * inv, my = sy * inv, mz = sz * inv;  // mean of d
    double out[9];
    out[0] = q0x + mx; out[1] = q0y + my; out[2] = q0z + mz;
    out[3] = fma(-sx, mx, xx); out[4] = fma(-sx, my, xy); out[5] = fma(-sx, mz, xz);
    out[6] = fma(-sy, my, yy); out[7] = fma(-sy, mz, yz); out[8] = fma(-sz, mz, zz);
    if (s >= 1) {
      const int64_t row = (int64_t)first_row + s - 1;
      count[row] = (uint32_t)cnt;
#pragma unroll
      for (int k = 0; k < 3; ++k) mean3[(int64_t)k * nvox + row] = out[k];
#pragma unroll
      for (int k = 0; k < 6; ++k) m2[(int64_t)k * nvox + row] = out[3 + k];
    } else {
      double* c = carry + (int64_t)blockIdx.x * 10;
#pragma unroll
      for (int k = 0; k < 9; ++k) c[k] = out[k];
      carry_cnt[blockIdx.x] = (uint32_t)cnt;
    }
  };
  // one thread per local segment; segments of >= kLongSeg points (large voxels) are queued for a warp each, and segments that are
  // long for THIS tile — at least kMidSeg points and twice the tile's mean — for a group of 8 lanes: a uniform cloud (10 +- 3 points
  // per voxel) keeps the balanced one-thread walk, a clustered one (most voxels 1-2 points, the terrain sheet 20-60) does not leave
  // 31 lanes waiting for the one that walks 40 points.  The threshold depends on the sorted keys only: results stay reproducible.
  const int mid_thr = max(kMidSeg, 2 * tile_n / (int)(nheads + 1));
  // L lanes per segment (L = 2, 4: clouds of many points per voxel — a tile of 1024 points then holds fewer segments than the CTA has
  // threads, and with one thread per segment half of the warps waited at the barrier below while the others walked ~10 points each):
  // lane l of the group sums points a + 1 + l, a + 1 + l + L, ..., folded by an xor butterfly inside the group (fixed order).
  for (unsigned s = threadIdx.x / L; s <= nheads; s += kThreads / L) {
    const int l = threadIdx.x & (L - 1);
    const int a = sm.seg_start[s], b = sm.seg_start[s + 1];
    if (b <= a) continue;  // only s == 0 can be empty (tile starts with a head)
    if (b - a >= kLongSeg) {
      if (l == 0) sm.long_seg[atomicAdd(&sm.n_long, 1u)] = (unsigned short)s;
      continue;
    }
    if (MID && b - a >= mid_thr) {
      if (l == 0) sm.mid_seg[atomicAdd(&sm.n_mid, 1u)] = (unsigned short)s;
      continue;
    }
    long long vx, vy, vz;
    unpack_key((unsigned long long)sm.seg_key[s], g, vx, vy, vz);
    const double cx = shift_center(vx, g.ox, g.vs), cy = shift_center(vy, g.oy, g.vs), cz = shift_center(vz, g.oz, g.vs);
    const double q0x = sm.x[sw(a)] - cx, q0y = sm.y[sw(a)] - cy, q0z = sm.z[sw(a)] - cz;
    double sx = 0, sy = 0, sz = 0, xx = 0, xy = 0, xz = 0, yy = 0, yz = 0, zz = 0;
    for (int p = a + 1 + l; p < b; p += L) {
      const int q = sw(p);
      const double dx = (sm.x[q] - cx) - q0x, dy = (sm.y[q] - cy) - q0y, dz = (sm.z[q] - cz) - q0z;
      sx += dx; sy += dy; sz += dz;
      xx = fma(dx, dx, xx); xy = fma(dx, dy, xy); xz = fma(dx, dz, xz);
      yy = fma(dy, dy, yy); yz = fma(dy, dz, yz); zz = fma(dz, dz, zz);
    }
    if (L > 1) {
      // the lanes of a group share s, a and b: they are here together (other groups of the warp may have left the loop)
      const unsigned gmask = ((1u << L) - 1u) << ((threadIdx.x & 31) & ~(L - 1));
#pragma unroll
      for (int o = L / 2; o > 0; o >>= 1) {
        sx += __shfl_xor_sync(gmask, sx, o); sy += __shfl_xor_sync(gmask, sy, o); sz += __shfl_xor_sync(gmask, sz, o);
        xx += __shfl_xor_sync(gmask, xx, o); xy += __shfl_xor_sync(gmask, xy, o); xz += __shfl_xor_sync(gmask, xz, o);
        yy += __shfl_xor_sync(gmask, yy, o); yz += __shfl_xor_sync(gmask, yz, o); zz += __shfl_xor_sync(gmask, zz, o);
      }
      if (l != 0) continue;
    }
    emit(s, b - a, q0x, q0y, q0z, sx, sy, sz, xx, xy, xz, yy, yz, zz);
  }
  __syncthreads();
  // middle segments (clustered clouds: a few voxels of the terrain sheet hold 20-60 points each while most voxels hold one or two —
  // walked by one thread each, such a voxel kept its whole warp waiting): a group of 8 lanes each, four per warp at a time; lane l of
  // the group sums points a + 1 + l, a + 1 + l + 8, ..., folded with a fixed xor butterfly inside the group
  if (MID) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, gl = lane & 7;
    const unsigned nm = sm.n_mid;
    // a warp takes four consecutive segments of the queue at a time (spreading the queue over all warps costs more issue slots than
    // the shorter tail wins: every warp then runs the butterfly for one or two active groups)
    for (unsigned i0 = warp * 4; i0 < nm; i0 += (kThreads / 32) * 4) {
      const unsigned i = i0 + (lane >> 3);
      const bool act = i < nm;
      const unsigned s = act ? sm.mid_seg[i] : 0u;
      const int a = act ? sm.seg_start[s] : 0, b = act ? sm.seg_start[s + 1] : 0;
      long long vx, vy, vz;
      unpack_key((unsigned long long)sm.seg_key[s], g, vx, vy, vz);
      const double cx = shift_center(vx, g.ox, g.vs), cy = shift_center(vy, g.oy, g.vs), cz = shift_center(vz, g.oz, g.vs);
      const double q0x = sm.x[sw(a)] - cx, q0y = sm.y[sw(a)] - cy, q0z = sm.z[sw(a)] - cz;
      double v[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
      for (int p = a + 1 + gl; p < b; p += 8) {
        const int q = sw(p);
        const double dx = (sm.x[q] - cx) - q0x, dy = (sm.y[q] - cy) - q0y, dz = (sm.z[q] - cz) - q0z;
        v[0] += dx; v[1] += dy; v[2] += dz;
        v[3] = fma(dx, dx, v[3]); v[4] = fma(dx, dy, v[4]); v[5] = fma(dx, dz, v[5]);
        v[6] = fma(dy, dy, v[6]); v[7] = fma(dy, dz, v[7]); v[8] = fma(dz, dz, v[8]);
      }
#pragma unroll
      for (int o = 4; o > 0; o >>= 1) {
#pragma unroll
        for (int k = 0; k < 9; ++k) v[k] += __shfl_xor_sync(0xffffffffu, v[k], o);
      }
      if (act && gl == 0) emit(s, b - a, q0x, q0y, q0z, v[0], v[1], v[2], v[3], v[4], v[5], v[6], v[7], v[8]);
    }
  }
  // long segments: lane l sums points a + 1 + l, a + 1 + l + 32, ...; the nine sums are folded with a fixed xor butterfly
  // (every lane ends with the same bits), so the result does not depend on scheduling
  {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const unsigned nl = sm.n_long;
    for (unsigned i = warp; i < nl; i += kThreads / 32) {
      const unsigned s = sm.long_seg[i];
      const int a = sm.seg_start[s], b = sm.seg_start[s + 1];
      long long vx, vy, vz;
      unpack_key((unsigned long long)sm.seg_key[s], g, vx, vy, vz);
      const double cx = shift_center(vx, g.ox, g.vs), cy = shift_center(vy, g.oy, g.vs), cz = shift_center(vz, g.oz, g.vs);
      const double q0x = sm.x[sw(a)] - cx, q0y = sm.y[sw(a)] - cy, q0z = sm.z[sw(a)] - cz;
      double v[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
      for (int p = a + 1 + lane; p < b; p += 32) {
        const int q = sw(p);
        const double dx = (sm.x[q] - cx) - q0x, dy = (sm.y[q] - cy) - q0y, dz = (sm.z[q] - cz) - q0z;
        v[0] += dx; v[1] += dy; v[2] += dz;
        v[3] = fma(dx, dx, v[3]); v[4] = fma(dx, dy, v[4]); v[5] = fma(dx, dz, v[5]);
        v[6] = fma(dy, dy, v[6]); v[7] = fma(dy, dz, v[7]); v[8] = fma(dz, dz, v[8]);
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
        for (int k = 0; k < 9; ++k) v[k] += __shfl_xor_sync(0xffffffffu, v[k], o);
      }
      if (lane == 0) emit(s, b - a, q0x, q0y, q0z, v[0], v[1], v[2], v[3], v[4], v[5], v[6], v[7], v[8]);
    }
  }
  // tiles that start with a head have no carry
  if (threadIdx.x == 0 && sm.seg_start[1] == 0 && nheads > 0) carry_cnt[blockIdx.x] = 0;
}

// Fix-up: a voxel continued over tiles t, t+1, ... (carry_cnt > 0 and the same row) gets its carries folded
// in, in tile order.  Almost every chain is a single tile (a voxel simply straddles one tile boundary): that
// case is one Chan update by one thread.  Longer chains (a voxel with more points than a tile) are handed to a
// warp each: lane l folds tiles t+l, t+l+32, ... sequentially and lane 0 folds the 32 lane results in lane
// order — a fixed association, so the result does not depend on scheduling.
__global__ void __launch_bounds__(kThreads) fixup_moments_kernel(const uint32_t* __restrict__ tile_first,
                                                                 const uint32_t* __restrict__ carry_cnt, const double* __restrict__ carry,
                                                                 int ntiles, int64_t nvox, uint32_t* __restrict__ count,
                                                                 double* __restrict__ mean3, double* __restrict__ m2,
                                                                 uint32_t* __restrict__ long_chains, uint32_t* __restrict__ n_long) {
  const int t = (int)((int64_t)blockIdx.x * blockDim.x + threadIdx.x);
  if (t >= ntiles) return;
  const uint32_t cc = carry_cnt[t];
  if (cc == 0) return;
  const uint32_t row = tile_first[t] - 1u;
  if (t > 0 && carry_cnt[t - 1] != 0 && tile_first[t - 1] - 1u == row) return;  // not the chain start
  if (t + 1 < ntiles && carry_cnt[t + 1] != 0 && tile_first[t + 1] - 1u == row) {  // chain of >= 2 tiles
    long_chains[atomicAdd(n_long, 1u)] = (uint32_t)t;
    return;
  }
  double na = (double)count[row], ma[3], Ma[6], mb[3], Mb[6];
  const double* c = carry + (int64_t)t * 10;
#pragma unroll
  for (int k = 0; k < 3; ++k) { ma[k] = mean3[(int64_t)k * nvox + row]; mb[k] = c[k]; }
#pragma unroll
  for (int k = 0; k < 6; ++k) { Ma[k] = m2[(int64_t)k * nvox + row]; Mb[k] = c[3 + k]; }
  chan_combine(na, ma, Ma, (double)cc, mb, Mb);
  count[row] = (uint32_t)na;
#pragma unroll
  for (int k = 0; k < 3; ++k) mean3[(int64_t)k * nvox + row] = ma[k];
#pragma unroll
  for (int k = 0; k < 6; ++k) m2[(int64_t)k * nvox + row] = Ma[k];
}

__global__ void __launch_bounds__(kThreads) fixup_long_chains_kernel(const uint32_t* __restrict__ tile_first,
                                                                     const uint32_t* __restrict__ carry_cnt, const double* __restrict__ carry,
                                                                     int ntiles, int64_t nvox, uint32_t* __restrict__ count,
                                                                     double* __restrict__ mean3, double* __restrict__ m2,
                                                                     const uint32_t* __restrict__ long_chains, const uint32_t* __restrict__ n_long) {
  const int lane = threadIdx.x & 31;
  const unsigned nchains = *n_long;
  for (unsigned w = (unsigned)(((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5); w < nchains; w += (gridDim.x * blockDim.x) >> 5) {
    const int t = (int)long_chains[w];
    const uint32_t row = tile_first[t] - 1u;
    double na = 0.0, ma[3] = {0, 0, 0}, Ma[6] = {0, 0, 0, 0, 0, 0};
    for (int u0 = t; u0 < ntiles; u0 += 32) {
      const int u = u0 + lane;
      const bool mine = u < ntiles && carry_cnt[u] != 0 && tile_first[u] - 1u == row;
      const unsigned ok = __ballot_sync(0xffffffffu, mine);
      const unsigned run = (ok == 0xffffffffu) ? 32u : (unsigned)(__ffs(~ok) - 1);  // leading tiles of the chain
      if ((unsigned)lane < run) {
        const double* c = carry + (int64_t)u * 10;
        const double mb[3] = {c[0], c[1], c[2]};
        const double Mb[6] = {c[3], c[4], c[5], c[6], c[7], c[8]};
        chan_combine(na, ma, Ma, (double)carry_cnt[u], mb, Mb);
      }
      if (run < 32u) break;
    }
    double nt = 0.0, mt[3] = {0, 0, 0}, Mt[6] = {0, 0, 0, 0, 0, 0};
    if (lane == 0) {
      nt = (double)count[row];
#pragma unroll
      for (int k = 0; k < 3; ++k) mt[k] = mean3[(int64_t)k * nvox + row];
#pragma unroll
      for (int k = 0; k < 6; ++k) Mt[k] = m2[(int64_t)k * nvox + row];
    }
    for (int l = 0; l < 32; ++l) {
      const double nb = __shfl_sync(0xffffffffu, na, l);
      double mb[3], Mb[6];
#pragma unroll
      for (int k = 0; k < 3; ++k) mb[k] = __shfl_sync(0xffffffffu, ma[k], l);
#pragma unroll
      for (int k = 0; k < 6; ++k) Mb[k] = __shfl_sync(0xffffffffu, Ma[k], l);
      if (lane == 0) chan_combine(nt, mt, Mt, nb, mb, Mb);
    }
    if (lane == 0) {
      count[row] = (uint32_t)nt;
#pragma unroll
      for (int k = 0; k < 3; ++k) mean3[(int64_t)k * nvox + row] = mt[k];
#pragma unroll
      for (int k = 0; k < 6; ++k) m2[(int64_t)k * nvox + row] = Mt[k];
    }
  }
}

// ---- point -> voxel map through a rank bitmap over the key space (dense key spaces) -------------------------
// Voxel rows are in key order, so the row of an occupied key is its RANK among the occupied keys.  The table holds, per
// 32 cells of the key space, {occupancy bits, row of the word's first occupied cell}: row(key) = first + popc(bits below
// the key's bit).  8 bytes per 32 cells = 2^total_bits / 4 bytes — 4 MB for the 2^24 cells of the benchmark grid against
// 64 MB for a row per cell (which fell out of the L2: 61 % hit rate, profiles/r1_d) — so the per-point lookups are cache
// hits and the kernel streams keys in, rows out.
template <typename KeyT>
__global__ void __launch_bounds__(kThreads) p2v_fill_kernel(const KeyT* __restrict__ voxel_keys, int64_t nvox, uint2* __restrict__ table) {
  const int lane = threadIdx.x & 31;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  const int64_t rounds = (nvox + stride - 1) / stride;
  for (int64_t it = 0; it < rounds; ++it) {  // whole warps iterate together (shuffles below)
    const int64_t v = it * stride + (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const bool valid = v < nvox;
    const KeyT k = valid ? voxel_keys[v] : (KeyT)0;
    const KeyT w = valid ? (KeyT)(k >> 5) : (KeyT)~(KeyT)0;  // no valid word is all ones
    unsigned bits = valid ? 1u << ((unsigned)k & 31u) : 0u;
    // the keys are sorted: the voxels of a word are neighbouring lanes.  Segmented OR scan, one atomic per run.
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const unsigned b2 = __shfl_up_sync(0xffffffffu, bits, o);
      const KeyT w2 = __shfl_up_sync(0xffffffffu, w, o);
      if (lane >= o && w2 == w) bits |= b2;
    }
    const KeyT wn = __shfl_down_sync(0xffffffffu, w, 1);
    if (valid && (lane == 31 || wn != w)) atomicOr(&table[w].x, bits);      // the occupancy words are zeroed by the caller
    if (valid && (v == 0 || (voxel_keys[v - 1] >> 5) != w)) table[w].y = (unsigned)v;  // the first voxel of a word names its row
  }
}
__device__ __forceinline__ uint32_t p2v_row(const uint2* __restrict__ table, unsigned long long key) {
  const uint2 t = __ldg(table + (key >> 5));
  return t.y + (uint32_t)__popc(t.x & ((1u << ((unsigned)key & 31u)) - 1u));
}
// Four keys per thread and iteration: what matters is how many lookups are in flight.
template <typename KeyT>
__global__ void __launch_bounds__(kThreads) p2v_lookup_kernel(const KeyT* __restrict__ keys, int64_t n, const uint2* __restrict__ table,
                                                              uint32_t* __restrict__ p2v) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  const bool vec = ((reinterpret_cast<uintptr_t>(keys) | reinterpret_cast<uintptr_t>(p2v)) & 15u) == 0 && sizeof(KeyT) == 4;
  const int64_t nvec = vec ? n / 4 : 0;
  for (int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; q < nvec; q += stride) {
    const uint4 k = __ldcs(reinterpret_cast<const uint4*>(keys) + q);
    const uint2 a = __ldg(table + (k.x >> 5)), b = __ldg(table + (k.y >> 5)), c = __ldg(table + (k.z >> 5)), d = __ldg(table + (k.w >> 5));
    uint4 r;
    r.x = a.y + (uint32_t)__popc(a.x & ((1u << (k.x & 31u)) - 1u));
    r.y = b.y + (uint32_t)__popc(b.x & ((1u << (k.y & 31u)) - 1u));
    r.z = c.y + (uint32_t)__popc(c.x & ((1u << (k.z & 31u)) - 1u));
    r.w = d.y + (uint32_t)__popc(d.x & ((1u << (k.w & 31u)) - 1u));
    __stcs(reinterpret_cast<uint4*>(p2v) + q, r);
  }
  for (int64_t i = nvec * 4 + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
    p2v[i] = p2v_row(table, (unsigned long long)__ldcs(keys + i));
}

// point2voxel[index[i]] = row[i].  Called with (index, row) pairs grouped by the leading bits of the index, so that the writes
// of a group stay inside a window of the output that the L2 holds and leave for HBM as whole lines.
__global__ void __launch_bounds__(kThreads) scatter_rows_kernel(const uint32_t* __restrict__ index, const uint32_t* __restrict__ row, int64_t n,
                                                                uint32_t* __restrict__ out) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) out[__ldcs(index + i)] = __ldcs(row + i);
}

// original index of every sorted position, read from the fourth slot of its record (8 of the record's 32 bytes = one sector)
__global__ void __launch_bounds__(kThreads) record_indices_kernel(const double* __restrict__ xyzw, const uint32_t* __restrict__ pos_sorted, int64_t n,
                                                                  uint32_t* __restrict__ idx_sorted) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; p < n; p += stride)
    idx_sorted[p] = (uint32_t)__double_as_longlong(__ldg(xyzw + 4 * (int64_t)__ldcs(pos_sorted + p) + 3));
}

// ---- closest to centroid ----------------------------------------------------------------------------
// distance = sqrt((X-cog_x)^2 + (Y-cog_y)^2 + (Z-cog_z)^2), numpy evaluation order, no FMA.
__device__ __forceinline__ double ref_distance(double px, double py, double pz, double cx, double cy, double cz) {
  const double dx = __dsub_rn(px, cx), dy = __dsub_rn(py, cy), dz = __dsub_rn(pz, cz);
  return __dsqrt_rn(__dadd_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)), __dmul_rn(dz, dz)));
}

template <typename KeyT>
__global__ void __launch_bounds__(kThreads) closest_kernel(const KeyT* __restrict__ keys, const uint32_t* __restrict__ pos_sorted,
                                                           int64_t n, const double4* __restrict__ xyzw, const double* __restrict__ cog3,
                                                           const uint32_t* __restrict__ tile_first, int64_t nvox,
                                                           double* __restrict__ min_dist, uint32_t* __restrict__ argmin_idx,
                                                           uint32_t* __restrict__ carry_cnt, double* __restrict__ carry) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  TileSmem<KeyT>& sm = *reinterpret_cast<TileSmem<KeyT>*>(smem_raw);
  __shared__ uint32_t idx_s[kSegTile];
  const int64_t base = (int64_t)blockIdx.x * kSegTile;
  KeyT key[kItems];
  bool valid[kItems], head[kItems];
  unsigned inc[kItems], nheads;
  int tile_n;
  const int64_t p0 = base + (int64_t)threadIdx.x * kItems;
#pragma unroll
  for (int j = 0; j < kItems; ++j) {
    const int p = threadIdx.x * kItems + j;
    if (p0 + j < n) {
      const double4 r = ld_record(xyzw + pos_sorted[p0 + j]);
      sm.x[sw(p)] = r.x; sm.y[sw(p)] = r.y; sm.z[sw(p)] = r.z;
      idx_s[p] = record_index(r);
    }
  }
  tile_prologue(sm, keys, n, base, key, valid, head, inc, nheads, tile_n);
  const uint32_t first_row = tile_first[blockIdx.x];
  if (threadIdx.x == 0) sm.n_long = 0;
  __syncthreads();
  auto emit = [&](unsigned s, int cnt, double best, uint32_t best_i) {
    if (s >= 1) {
      const int64_t row = (int64_t)first_row + s - 1;
      min_dist[row] = best;
      argmin_idx[row] = best_i;
    } else {
      double* c = carry + (int64_t)blockIdx.x * 10;
      c[0] = best;
      c[1] = __longlong_as_double((long long)best_i);
      carry_cnt[blockIdx.x] = (uint32_t)cnt;
    }
  };
  for (unsigned s = threadIdx.x; s <= nheads; s += kThreads) {
    const int a = sm.seg_start[s], b = sm.seg_start[s + 1];
    if (b <= a) continue;
    if (b - a >= kLongSeg) {  // large voxels: a warp each, below
      sm.long_seg[atomicAdd(&sm.n_long, 1u)] = (unsigned short)s;
      continue;
    }
    const int64_t row = (int64_t)first_row + s - 1;  // s == 0: first_row - 1 (continued voxel)
    const double cx = cog3[row], cy = cog3[nvox + row], cz = cog3[2 * nvox + row];
    double best = INFINITY;
    uint32_t best_i = 0xffffffffu;
    for (int p = a; p < b; ++p) {
      const double d = ref_distance(sm.x[sw(p)], sm.y[sw(p)], sm.z[sw(p)], cx, cy, cz);
      if (d < best || best_i == 0xffffffffu) { best = d; best_i = idx_s[p]; }  // strict <: first (lowest index) wins ties
    }
    emit(s, b - a, best, best_i);
  }
  __syncthreads();
  {
    // lexicographic minimum of (distance, tile position) over the lanes: the same winner as the sequential walk
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const unsigned nl = sm.n_long;
    for (unsigned i = warp; i < nl; i += kThreads / 32) {
      const unsigned s = sm.long_seg[i];
      const int a = sm.seg_start[s], b = sm.seg_start[s + 1];
      const int64_t row = (int64_t)first_row + s - 1;
      const double cx = cog3[row], cy = cog3[nvox + row], cz = cog3[2 * nvox + row];
      double best = INFINITY;
      int best_p = 0x7fffffff;
      for (int p = a + lane; p < b; p += 32) {
        const double d = ref_distance(sm.x[sw(p)], sm.y[sw(p)], sm.z[sw(p)], cx, cy, cz);
        if (d < best || best_p == 0x7fffffff) { best = d; best_p = p; }
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        const double od = __shfl_xor_sync(0xffffffffu, best, o);
        const int op = __shfl_xor_sync(0xffffffffu, best_p, o);
        if (op != 0x7fffffff && (best_p == 0x7fffffff || od < best || (od == best && op < best_p))) { best = od; best_p = op; }
      }
      if (lane == 0) emit(s, b - a, best, idx_s[best_p]);
    }
  }
  if (threadIdx.x == 0 && sm.seg_start[1] == 0 && nheads > 0) carry_cnt[blockIdx.x] = 0;
}

__global__ void __launch_bounds__(kThreads) fixup_closest_kernel(const uint32_t* __restrict__ tile_first,
                                                                 const uint32_t* __restrict__ carry_cnt, const double* __restrict__ carry,
                                                                 int ntiles, double* __restrict__ min_dist, uint32_t* __restrict__ argmin_idx) {
  const int t = (int)((int64_t)blockIdx.x * blockDim.x + threadIdx.x);
  if (t >= ntiles || carry_cnt[t] == 0) return;
  const uint32_t row = tile_first[t] - 1u;
  if (t > 0 && carry_cnt[t - 1] != 0 && tile_first[t - 1] - 1u == row) return;
  double best = min_dist[row];
  uint32_t best_i = argmin_idx[row];
  for (int u = t; u < ntiles && carry_cnt[u] != 0 && tile_first[u] - 1u == row; ++u) {
    const double d = carry[(int64_t)u * 10];
    if (d < best) { best = d; best_i = (uint32_t)__double_as_longlong(carry[(int64_t)u * 10 + 1]); }
  }
  min_dist[row] = best;
  argmin_idx[row] = best_i;
}

// ---- merge of partial voxel rows (multi-GPU) ----------------------------------------------------------
// Rows with equal keys are adjacent (keys sorted, stable, so source-rank order is kept); order[p] is
// the source row of sorted position p.  A key has at most one partial row per source rank, so the
// thread at each run head simply walks its run in global memory (runs may cross tiles freely).
// Sources of the partial rows.  SOA: columns count_in / mean3_in / m2_in (column stride in_stride).  ROWS: row-major
// {key, count, mean[3], M2[6]} of 11 x 8 bytes as exchanged between GPUs (`rows`).  MIXED: source index < n_local is row
// local_first + index of the SOA columns (this GPU's own rows, never copied), the others are `rows`[index - n_local].
enum { kSrcSoa = 0, kSrcRows = 1, kSrcMixed = 2 };
template <typename KeyT, int SRC>
__global__ void __launch_bounds__(kThreads) merge_partials_kernel(const KeyT* __restrict__ keys, const uint32_t* __restrict__ order,
                                                                  int64_t n, const uint32_t* __restrict__ count_in,
                                                                  const double* __restrict__ mean3_in, const double* __restrict__ m2_in,
                                                                  int64_t in_stride, const double* __restrict__ rows, int64_t n_local,
                                                                  int64_t local_first, const uint32_t* __restrict__ tile_first, int64_t nvox,
                                                                  KeyT* __restrict__ voxel_keys, uint32_t* __restrict__ count,
                                                                  double* __restrict__ mean3, double* __restrict__ m2) {
  __shared__ unsigned scan_s[33];
  const int64_t base = (int64_t)blockIdx.x * kSegTile;
  const int64_t p0 = base + (int64_t)threadIdx.x * kItems;
  KeyT key[kItems], prev;
  bool valid[kItems], head[kItems];
  load_tile_keys(keys, n, base, key, prev, valid);
  const unsigned local = head_flags(key, prev, valid, p0, head);
  unsigned nheads;
  unsigned run = block_exclusive_scan(local, scan_s, &nheads);
  const uint32_t first_row = tile_first[blockIdx.x];
#pragma unroll
  for (int j = 0; j < kItems; ++j) {
    if (!head[j]) continue;
    const int64_t row = (int64_t)first_row + run;
    ++run;
    double na = 0.0, ma[3] = {0, 0, 0}, Ma[6] = {0, 0, 0, 0, 0, 0};
    int64_t q = p0 + j;
    do {
      int64_t src = order[q];
      double mb[3], Mb[6], nb;
      const bool from_rows = SRC == kSrcRows || (SRC == kSrcMixed && src >= n_local);
      if (SRC == kSrcMixed) src = from_rows ? src - n_local : src + local_first;
      if (from_rows) {
        const double* row = rows + src * kPartialRowWords;
        nb = (double)__double_as_longlong(row[1]);
#pragma unroll
        for (int k = 0; k < 3; ++k) mb[k] = row[2 + k];
#pragma unroll
        for (int k = 0; k < 6; ++k) Mb[k] = row[5 + k];
      } else {
        nb = (double)count_in[src];
#pragma unroll
        for (int k = 0; k < 3; ++k) mb[k] = mean3_in[(int64_t)k * in_stride + src];
#pragma unroll
        for (int k = 0; k < 6; ++k) Mb[k] = m2_in[(int64_t)k * in_stride + src];
      }
      chan_combine(na, ma, Ma, nb, mb, Mb);
      ++q;
    } while (q < n && keys[q] == key[j]);
    voxel_keys[row] = key[j];
    count[row] = (uint32_t)na;
#pragma unroll
    for (int k = 0; k < 3; ++k) mean3[(int64_t)k * nvox + row] = ma[k];
#pragma unroll
    for (int k = 0; k < 6; ++k) m2[(int64_t)k * nvox + row] = Ma[k];
  }
}

// ---- merge of a rank's OWN partial rows with the (few) rows it received, without re-sorting the own rows ---------------------
// Both tables are key-sorted with unique keys: `own` = rows [own_first, own_first + n_own) of the local table (global keys given
// separately), `u` = the received rows, sorted and combined per key beforehand.  u_new_before[j] = number of u rows in front of j
// whose key the own rows do NOT hold (exclusive prefix, [nu + 1]); the merged table has n_own + u_new_before[nu] rows.
//   own row i whose key u does not hold -> output row i + u_new_before[lower_bound(u, key_i)]        (merge_own_rows_kernel: the bulk)
//   u row j -> output row lower_bound(own, key_j) + u_new_before[j]; combined (Chan: own first) with the own row of the same key if
//   there is one                                                                                   (merge_u_rows_kernel: few rows)
// One pass over the own rows (84 B read + written per row) instead of a radix sort of all keys + a gather-merge.
__device__ __forceinline__ int64_t lower_bound_u64(const unsigned long long* __restrict__ keys, int64_t n, unsigned long long k) {
  int64_t lo = 0, hi = n;
  while (lo < hi) {
    const int64_t mid = (lo + hi) >> 1;
    if (__ldg(keys + mid) < k) lo = mid + 1; else hi = mid;
  }
  return lo;
}
constexpr int kMergeChunks = 4;
template <typename KeyT>
__global__ void __launch_bounds__(kThreads) merge_own_rows_kernel(const unsigned long long* __restrict__ own_keys, const uint32_t* __restrict__ own_count,
                                                                  const double* __restrict__ own_mean3, const double* __restrict__ own_m2,
                                                                  int64_t own_stride, int64_t own_first, int64_t n_own,
                                                                  const unsigned long long* __restrict__ u_keys, const uint32_t* __restrict__ u_count,
                                                                  const double* __restrict__ u_mean3, const double* __restrict__ u_m2, int64_t nu,
                                                                  const int64_t* __restrict__ u_new_before, int64_t nvox, KeyT* __restrict__ out_keys,
                                                                  uint32_t* __restrict__ out_count, double* __restrict__ out_mean3,
                                                                  double* __restrict__ out_m2, Grid g, StatsOut so, int with_stats) {
  // The own keys ascend with the thread, so do their positions in u: two threads of the CTA bracket the CTA's slice of u
  // (a full binary search per row was a chain of ~18 dependent L2 reads in front of every row: 1.2 ms for 3.9e7 rows,
  // twice what its 84 B per row cost), every row then searches only inside that bracket (usually 0-2 keys).
  // A CTA takes kMergeChunks x 256 consecutive rows behind ONE bracket search (~18 dependent L2 reads by two edge threads), and its
  // first 256 rows — which do not depend on the bracket — are requested into L2 before it (prefetches: no registers held over the
  // barrier, the kernel stays as lean as the plain statistics kernel).  (One search per 256 rows with the loads behind the barrier:
  // every CTA idled ~5 us before it asked for its first byte, 2.59 ms for 3.9e7 rows against 2.02 ms of the statistics alone.)
  __shared__ int64_t bracket[2];
  const int64_t c0 = (int64_t)blockIdx.x * (kThreads * kMergeChunks);
  const int64_t c1 = min(n_own, c0 + (int64_t)kThreads * kMergeChunks);
  if (c0 + threadIdx.x < c1) {
    const int64_t i = c0 + threadIdx.x, src = own_first + i;
    prefetch_l2(own_keys + i);
    prefetch_l2(own_count + src);
#pragma unroll
    for (int k = 0; k < 3; ++k) prefetch_l2(own_mean3 + (int64_t)k * own_stride + src);
#pragma unroll
    for (int k = 0; k < 6; ++k) prefetch_l2(own_m2 + (int64_t)k * own_stride + src);
  }
  if (nu > 0 && threadIdx.x < 2) bracket[threadIdx.x] = lower_bound_u64(u_keys, nu, own_keys[threadIdx.x == 0 ? c0 : c1 - 1]);
  __syncthreads();
  for (int64_t i = c0 + threadIdx.x; i < c1; i += kThreads) {
    const int64_t src = own_first + i;
    const unsigned long long key = own_keys[i];
    const unsigned na = own_count[src];
    double ma[3], Ma[6];
#pragma unroll
    for (int k = 0; k < 3; ++k) ma[k] = own_mean3[(int64_t)k * own_stride + src];
#pragma unroll
    for (int k = 0; k < 6; ++k) Ma[k] = own_m2[(int64_t)k * own_stride + src];
    int64_t row = i;
    if (nu > 0) {
      const int64_t r = bracket[0] + lower_bound_u64(u_keys + bracket[0], bracket[1] - bracket[0], key);
      if (r < nu && __ldg(u_keys + r) == key) continue;  // u holds the key too: merge_u_rows_kernel combines the two rows (keeps this kernel lean)
      row += u_new_before[r];
    }
    out_keys[row] = (KeyT)key;
    out_count[row] = na;
    if (out_mean3) {
#pragma unroll
      for (int k = 0; k < 3; ++k) out_mean3[(int64_t)k * nvox + row] = ma[k];
    }
    if (out_m2) {
#pragma unroll
      for (int k = 0; k < 6; ++k) out_m2[(int64_t)k * nvox + row] = Ma[k];
    }
    if (with_stats) emit_voxel_stats(key, na, ma, Ma, g, row, so);
  }
}
template <typename KeyT>
__global__ void __launch_bounds__(kThreads) merge_u_rows_kernel(const unsigned long long* __restrict__ own_keys, const uint32_t* __restrict__ own_count,
                                                                const double* __restrict__ own_mean3, const double* __restrict__ own_m2,
                                                                int64_t own_stride, int64_t own_first, int64_t n_own,
                                                                  const unsigned long long* __restrict__ u_keys, const uint32_t* __restrict__ u_count,
                                                                  const double* __restrict__ u_mean3, const double* __restrict__ u_m2, int64_t nu,
                                                                  const int64_t* __restrict__ u_new_before, int64_t nvox, KeyT* __restrict__ out_keys,
                                                                  uint32_t* __restrict__ out_count, double* __restrict__ out_mean3,
                                                                  double* __restrict__ out_m2, Grid g, StatsOut so, int with_stats) {
  const int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= nu) return;
  const unsigned long long key = u_keys[j];
  const int64_t at = lower_bound_u64(own_keys, n_own, key);
  const int64_t row = at + u_new_before[j];
  double nb = (double)u_count[j], mb[3], Mb[6];
#pragma unroll
  for (int k = 0; k < 3; ++k) mb[k] = u_mean3[(int64_t)k * nu + j];
#pragma unroll
  for (int k = 0; k < 6; ++k) Mb[k] = u_m2[(int64_t)k * nu + j];
  if (u_new_before[j + 1] == u_new_before[j]) {  // the own rows hold this key: own row first, then the received one (Chan)
    const int64_t src = own_first + at;
    double na = (double)own_count[src], ma[3], Ma[6];
#pragma unroll
    for (int k = 0; k < 3; ++k) ma[k] = own_mean3[(int64_t)k * own_stride + src];
#pragma unroll
    for (int k = 0; k < 6; ++k) Ma[k] = own_m2[(int64_t)k * own_stride + src];
    chan_combine(na, ma, Ma, nb, mb, Mb);
    nb = na;
#pragma unroll
    for (int k = 0; k < 3; ++k) mb[k] = ma[k];
#pragma unroll
    for (int k = 0; k < 6; ++k) Mb[k] = Ma[k];
  }
  out_keys[row] = (KeyT)key;
  out_count[row] = (uint32_t)nb;
  if (out_mean3) {
#pragma unroll
    for (int k = 0; k < 3; ++k) out_mean3[(int64_t)k * nvox + row] = mb[k];
  }
  if (out_m2) {
#pragma unroll
    for (int k = 0; k < 6; ++k) out_m2[(int64_t)k * nvox + row] = Mb[k];
  }
  if (with_stats) emit_voxel_stats(key, (unsigned)nb, mb, Mb, g, row, so);
}

// local voxel keys -> keys in the GLOBAL layout (unpack with the local grid, repack with the global one) + the histogram of
// their leading bits.  The table is key-sorted, so neighbouring rows share a bin: one atomic per run of equal bins in a warp.
template <typename KeyT>
__global__ void __launch_bounds__(kThreads) global_keys_kernel(const KeyT* __restrict__ voxel_keys, int64_t nvox, Grid lg, Grid gg, int hist_shift,
                                                               unsigned long long* __restrict__ keys_out, uint32_t* __restrict__ hist) {
  // Every warp takes a contiguous range of rows.  The keys are sorted, so a bin is a run of thousands of rows: the warp carries
  // (bin, count) of its current run from round to round and touches the histogram only when the bin changes (one atomic per warp and
  // bin instead of one per warp and round: 1.2e6 atomics on a handful of addresses were 0.15 of this kernel's 0.23 ms at 3.9e7 rows).
  const int lane = threadIdx.x & 31;
  const int64_t warps = ((int64_t)gridDim.x * blockDim.x) >> 5, wid = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t per = ((nvox + warps - 1) / warps + 31) / 32 * 32;
  const int64_t v0 = min(nvox, wid * per), v1 = min(nvox, v0 + per);
  unsigned cur_bin = 0xffffffffu, cur_cnt = 0;  // the same in all lanes
  for (int64_t base = v0; base < v1; base += 32) {
    const int64_t v = base + lane;
    const bool valid = v < v1;
    unsigned bin = 0xffffffffu;
    if (valid) {
      long long vx, vy, vz;
      unpack_key((unsigned long long)voxel_keys[v], lg, vx, vy, vz);
      const int sxy = gg.by + gg.bz;
      unsigned long long k = (unsigned long long)(vz - gg.minz);
      if (gg.bz < 64) k |= (unsigned long long)(vy - gg.miny) << gg.bz;
      if (sxy < 64) k |= (unsigned long long)(vx - gg.minx) << sxy;
      keys_out[v] = k;
      bin = (unsigned)(k >> hist_shift);
    }
    const unsigned prev = __shfl_up_sync(0xffffffffu, bin, 1);
    const bool head = valid && (lane == 0 || prev != bin);
    const unsigned heads = __ballot_sync(0xffffffffu, head);
    const int nvalid = __popc(__ballot_sync(0xffffffffu, valid));
    const unsigned first_bin = __shfl_sync(0xffffffffu, bin, 0);
    const unsigned after0 = heads & ~1u;  // heads behind lane 0: the round holds more than one run
    const int end0 = after0 ? __ffs(after0) - 1 : nvalid;
    if (first_bin == cur_bin) {
      cur_cnt += (unsigned)end0;
    } else {
      if (cur_cnt && lane == 0) atomicAdd(hist + cur_bin, cur_cnt);
      cur_bin = first_bin;
      cur_cnt = (unsigned)end0;
    }
    if (after0) {
      if (lane == 0) atomicAdd(hist + cur_bin, cur_cnt);
      const int last_head = 31 - __clz(heads);
      if (head && lane != 0 && lane != last_head) {
        const unsigned after = heads & ~((2u << lane) - 1u);
        atomicAdd(hist + bin, (unsigned)(__ffs(after) - 1 - lane));
      }
      cur_bin = __shfl_sync(0xffffffffu, bin, last_head);
      cur_cnt = (unsigned)(nvalid - last_head);
    }
  }
  if (cur_cnt && lane == 0) atomicAdd(hist + cur_bin, cur_cnt);
}

// rows [begin, end) of a partial table (SOA, column stride nvox) -> exchange rows (row-major, 11 x 8 bytes) at rows_out.
// rows_out may be PEER memory (another GPU's receive buffer over NVLink): a CTA transposes 256 rows through shared memory
// and writes them out as one contiguous 22 KB run of consecutive 8-byte stores (row-strided 8-byte remote stores ran at
// 76 GB/s, a twelfth of the link).
__global__ void __launch_bounds__(kThreads) pack_partial_rows_kernel(const unsigned long long* __restrict__ gkeys, const uint32_t* __restrict__ count,
                                                                     const double* __restrict__ mean3, const double* __restrict__ m2, int64_t nvox,
                                                                     int64_t begin, int64_t end, double* __restrict__ rows_out) {
  __shared__ double tile[kThreads * kPartialRowWords];
  for (int64_t base = begin + (int64_t)blockIdx.x * kThreads; base < end; base += (int64_t)gridDim.x * kThreads) {
    const int rows = (int)min((int64_t)kThreads, end - base);
    const int64_t v = base + threadIdx.x;
    if ((int)threadIdx.x < rows) {
      double* row = tile + threadIdx.x * kPartialRowWords;  // 11 words per row: odd stride, conflict-free
      row[0] = __longlong_as_double((long long)gkeys[v]);
      row[1] = __longlong_as_double((long long)count[v]);
#pragma unroll
      for (int k = 0; k < 3; ++k) row[2 + k] = mean3[(int64_t)k * nvox + v];
#pragma unroll
      for (int k = 0; k < 6; ++k) row[5 + k] = m2[(int64_t)k * nvox + v];
    }
    __syncthreads();
    double* out = rows_out + (base - begin) * kPartialRowWords;
    for (int i = threadIdx.x; i < rows * kPartialRowWords; i += kThreads) out[i] = tile[i];
    __syncthreads();
  }
}

template <typename KeyT>
size_t tile_smem_bytes() { return sizeof(TileSmem<KeyT>); }
}  // namespace

// ------------------------------------------------------------------------------------------
// C ABI
// ------------------------------------------------------------------------------------------
extern "C" size_t vapc_segment_workspace_bytes(int64_t n) { return seg_ws_bytes(n); }

extern "C" int vapc_count_voxels(const void* keys_sorted, int64_t n, int32_t key_bytes, int64_t* nvoxels, void* ws,
                                 size_t ws_bytes, void* stream) {
  if (n < 0 || !nvoxels) return vapc_set_error(VAPC_E_INVALID, "vapc_count_voxels: bad arguments");
  cudaStream_t s = as_stream(stream);
  if (n == 0) {
    VAPC_RETURN_IF_CUDA(cudaMemsetAsync(nvoxels, 0, sizeof(int64_t), s));
    return VAPC_OK;
  }
  if (n >= (int64_t)1 << 32) return vapc_set_error(VAPC_E_INVALID, "vapc_count_voxels: n must be < 2^32 per GPU");
  if (!ws || ws_bytes < seg_ws_bytes(n)) return vapc_set_error(VAPC_E_WORKSPACE, "vapc_count_voxels: workspace too small");
  if (key_bytes != 4 && key_bytes != 8) return vapc_set_error(VAPC_E_INVALID, "vapc_count_voxels: key_bytes must be 4 or 8");
  if ((uintptr_t)keys_sorted & 15u) return vapc_set_error(VAPC_E_INVALID, "vapc_count_voxels: keys must be 16-byte aligned");
  const SegWs w = carve(ws, n);
  const int64_t nt = seg_tiles(n);
  if (key_bytes == 4)
    VAPC_LAUNCH(count_heads_kernel<uint32_t>, (unsigned)nt, kThreads, 0, s, static_cast<const uint32_t*>(keys_sorted), n, w.tile_heads);
  else
    VAPC_LAUNCH(count_heads_kernel<unsigned long long>, (unsigned)nt, kThreads, 0, s, static_cast<const unsigned long long*>(keys_sorted), n, w.tile_heads);
  VAPC_RETURN_IF_CUDA(scan_u32_exclusive(w.tile_heads, w.tile_heads, nt, w.total, w.scan_ws, s));
  VAPC_RETURN_IF_CUDA(cudaMemcpyAsync(nvoxels, w.total, sizeof(int64_t), cudaMemcpyDeviceToDevice, s));
  return VAPC_OK;
}

static int g_reduce_lanes = 0;  // 0: from n / V; 1, 2, 4: forced (tests, measurements)
extern "C" int vapc_reduce_debug_lanes(int32_t lanes) {
  g_reduce_lanes = (lanes == 1 || lanes == 2 || lanes == 4) ? lanes : 0;
  return VAPC_OK;
}

extern "C" int vapc_reduce_moments(const void* keys_sorted, const uint32_t* pos_sorted, int64_t n, const vapc_grid_t* grid_host,
                                   const double* xyzw, int64_t nvoxels_host, void* voxel_keys, uint32_t* start, uint32_t* count,
                                   uint32_t* first_idx, double* mean3, double* m2, uint32_t* row_sorted, uint32_t* idx_sorted,
                                   void* ws, size_t ws_bytes, void* stream) {
  if (!grid_host || n < 0 || nvoxels_host < 0) return vapc_set_error(VAPC_E_INVALID, "vapc_reduce_moments: bad arguments");
  if (n == 0) return VAPC_OK;
  if (!keys_sorted || !pos_sorted || !xyzw || !voxel_keys || !start || !count || !first_idx || !mean3 || !m2)
    return vapc_set_error(VAPC_E_INVALID, "vapc_reduce_moments: null buffer");
  if (!ws || ws_bytes < seg_ws_bytes(n)) return vapc_set_error(VAPC_E_WORKSPACE, "vapc_reduce_moments: workspace too small");
  if (((uintptr_t)keys_sorted | (uintptr_t)pos_sorted | (uintptr_t)idx_sorted | (uintptr_t)row_sorted) & 15u || ((uintptr_t)xyzw & 31u))
    return vapc_set_error(VAPC_E_INVALID, "vapc_reduce_moments: keys / positions / indices must be 16-byte and xyzw 32-byte aligned");
  const SegWs w = carve(ws, n);
  const Grid g = to_device_grid(grid_host);
  const int64_t nt = seg_tiles(n);
  cudaStream_t s = as_stream(stream);
  const double4* rec = reinterpret_cast<const double4*>(xyzw);
  // the instantiation follows the mean number of points per voxel (a function of n and V only: results stay reproducible)
  const bool mid = n < (int64_t)VAPC_MID_MEAN * nvoxels_host;  // few points per voxel on average: see reduce_moments_kernel
  // measured at 1e8 points: 10 points per voxel 1.077 / 1.127 / 1.320 ms with 1 / 2 / 4 lanes, 80 points per voxel 1.584 / 1.511 / 1.472 ms
  const int lanes = g_reduce_lanes > 0 ? g_reduce_lanes : (n >= 32 * nvoxels_host ? 4 : 1);
#define VAPC_REDUCE(K, MID, L)                                                                                                               \
  do {                                                                                                                                       \
    const size_t smem = tile_smem_bytes<K>();                                                                                                \
    VAPC_RETURN_IF_CUDA(cudaFuncSetAttribute((reduce_moments_kernel<K, MID, L>), cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));   \
    VAPC_LAUNCH((reduce_moments_kernel<K, MID, L>), (unsigned)nt, kThreads, smem, s, static_cast<const K*>(keys_sorted), pos_sorted, n, g,   \
                rec, w.tile_heads, nvoxels_host, static_cast<K*>(voxel_keys), start, count, first_idx, mean3, m2, row_sorted, idx_sorted,    \
                w.carry_cnt, w.carry);                                                                                                       \
  } while (0)
#define VAPC_REDUCE_K(K)                                  \
  do {                                                    \
    if (mid) VAPC_REDUCE(K, true, 1);                     \
    else if (lanes == 4) VAPC_REDUCE(K, false, 4);        \
    else if (lanes == 2) VAPC_REDUCE(K, false, 2);        \
    else VAPC_REDUCE(K, false, 1);                        \
  } while (0)
  if (grid_host->key_bytes == 4) VAPC_REDUCE_K(uint32_t); else VAPC_REDUCE_K(unsigned long long);
#undef VAPC_REDUCE_K
#undef VAPC_REDUCE
  VAPC_RETURN_IF_CUDA(cudaMemsetAsync(w.n_long, 0, sizeof(uint32_t), s));
  VAPC_LAUNCH(fixup_moments_kernel, (unsigned)((nt + kThreads - 1) / kThreads), kThreads, 0, s, w.tile_heads, w.carry_cnt, w.carry, (int)nt,
              nvoxels_host, count, mean3, m2, w.long_chains, w.n_long);
  VAPC_LAUNCH(fixup_long_chains_kernel, kNumSMs, kThreads, 0, s, w.tile_heads, w.carry_cnt, w.carry, (int)nt, nvoxels_host, count, mean3, m2,
              w.long_chains, w.n_long);
  VAPC_CHECK_LAUNCH();
  return VAPC_OK;
}

extern "C" int vapc_record_indices(const double* xyzw, const uint32_t* pos_sorted, int64_t n, uint32_t* idx_sorted, void* stream) {
  if (n < 0) return vapc_set_error(VAPC_E_INVALID, "vapc_record_indices: n < 0");
  if (n == 0) return VAPC_OK;
  if (!xyzw || !pos_sorted || !idx_sorted) return vapc_set_error(VAPC_E_INVALID, "vapc_record_indices: null buffer");
  const int grid = (int)std::min<int64_t>((n + kThreads * 4 - 1) / (kThreads * 4), (int64_t)kNumSMs * 16);
  VAPC_LAUNCH(record_indices_kernel, grid, kThreads, 0, as_stream(stream), xyzw, pos_sorted, n, idx_sorted);
  VAPC_CHECK_LAUNCH();
  return VAPC_OK;
}

extern "C" int vapc_scatter_u32(const uint32_t* index, const uint32_t* values, int64_t n, uint32_t* out, void* stream) {
  if (n < 0) return vapc_set_error(VAPC_E_INVALID, "vapc_scatter_u32: n < 0");
  if (n == 0) return VAPC_OK;
  if (!index || !values || !out) return vapc_set_error(VAPC_E_INVALID, "vapc_scatter_u32: null buffer");
  const int grid = (int)std::min<int64_t>((n + kThreads * 4 - 1) / (kThreads * 4), (int64_t)kNumSMs * 16);
  VAPC_LAUNCH(scatter_rows_kernel, grid, kThreads, 0, as_stream(stream), index, values, n, out);
  VAPC_CHECK_LAUNCH();
  return VAPC_OK;
}

extern "C" size_t vapc_point2voxel_table_bytes(int32_t total_bits) {
  return total_bits < 0 || total_bits > 32 ? 0 : (sizeof(uint2) << (total_bits > 5 ? total_bits - 5 : 0));
}

// CTAs per SM of p2v_lookup_kernel (a grid-stride kernel).  A caller that runs the lookup on a side stream next to another kernel
// lowers it so that the lookup leaves threads and registers of every SM to its neighbour.
static int g_p2v_ctas_per_sm = 16;
extern "C" int vapc_point2voxel_set_ctas_per_sm(int32_t ctas) {
  if (ctas < 1 || ctas > 32) return vapc_set_error(VAPC_E_INVALID, "vapc_point2voxel_set_ctas_per_sm: 1..32");
  g_p2v_ctas_per_sm = ctas;
  return VAPC_OK;
}

extern "C" int vapc_point2voxel_dense(const void* keys_unsorted, int64_t n, const void* voxel_keys, int64_t nvoxels, int32_t key_bytes,
                                      int32_t total_bits, void* table, size_t table_bytes, uint32_t* point2voxel, void* stream) {
  if (n < 0 || nvoxels < 0 || total_bits < 0 || total_bits > 32) return vapc_set_error(VAPC_E_INVALID, "vapc_point2voxel_dense: bad arguments");
  if (!table || table_bytes < vapc_point2voxel_table_bytes(total_bits)) return vapc_set_error(VAPC_E_WORKSPACE, "vapc_point2voxel_dense: table too small");
  if (n == 0) return VAPC_OK;
  cudaStream_t s = as_stream(stream);
  const int g1 = (int)std::min<int64_t>((nvoxels + kThreads - 1) / kThreads, (int64_t)kNumSMs * 16);
  const int g2 = (int)std::min<int64_t>((n + kThreads * 4 - 1) / (kThreads * 4), (int64_t)kNumSMs * g_p2v_ctas_per_sm);
  VAPC_RETURN_IF_CUDA(cudaMemsetAsync(table, 0, vapc_point2voxel_table_bytes(total_bits), s));
  if (key_bytes == 4) {
    VAPC_LAUNCH(p2v_fill_kernel<uint32_t>, g1 < 1 ? 1 : g1, kThreads, 0, s, static_cast<const uint32_t*>(voxel_keys), nvoxels, static_cast<uint2*>(table));
    VAPC_LAUNCH(p2v_lookup_kernel<uint32_t>, g2 < 1 ? 1 : g2, kThreads, 0, s, static_cast<const uint32_t*>(keys_unsorted), n,
                static_cast<const uint2*>(table), point2voxel);
  } else if (key_bytes == 8) {
    typedef unsigned long long K;
    VAPC_LAUNCH(p2v_fill_kernel<K>, g1 < 1 ? 1 : g1, kThreads, 0, s, static_cast<const K*>(voxel_keys), nvoxels, static_cast<uint2*>(table));
    VAPC_LAUNCH(p2v_lookup_kernel<K>, g2 < 1 ? 1 : g2, kThreads, 0, s, static_cast<const K*>(keys_unsorted), n, static_cast<const uint2*>(table),
                point2voxel);
  } else {
    return vapc_set_error(VAPC_E_INVALID, "vapc_point2voxel_dense: key_bytes must be 4 or 8");
  }
  VAPC_CHECK_LAUNCH();
  return VAPC_OK;
}

extern "C" int vapc_run_lengths(const uint64_t* keys_sorted, const uint32_t* weights, int64_t n, int64_t nruns_host, uint64_t* run_keys,
                                int64_t* run_weight, void* ws, size_t ws_bytes, void* stream) {
  if (n < 0 || nruns_host < 0) return vapc_set_error(VAPC_E_INVALID, "vapc_run_lengths: bad arguments");
  if (n == 0) return VAPC_OK;
  if (!keys_sorted || !run_keys || !run_weight) return vapc_set_error(VAPC_E_INVALID, "vapc_run_lengths: null buffer");
  if (!ws || ws_bytes < seg_ws_bytes(n)) return vapc_set_error(VAPC_E_WORKSPACE, "vapc_run_lengths: workspace too small");
  const SegWs w = carve(ws, n);
  cudaStream_t s = as_stream(stream);
  VAPC_RETURN_IF_CUDA(cudaMemsetAsync(run_weight, 0, (size_t)nruns_host * sizeof(int64_t), s));
  typedef unsigned long long K;
  const size_t smem = tile_smem_bytes<K>();
  VAPC_RETURN_IF_CUDA(cudaFuncSetAttribute(run_lengths_kernel<K>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  VAPC_LAUNCH(run_lengths_kernel<K>, (unsigned)seg_tiles(n), kThreads, smem, s, reinterpret_cast<const K*>(keys_sorted), weights, n, w.tile_heads,
              reinterpret_cast<K*>(run_keys), reinterpret_cast<unsigned long long*>(run_weight));
  VAPC_CHECK_LAUNCH();
  return VAPC_OK;
}

extern "C" int vapc_closest_to_cog(const void* keys_sorted, const uint32_t* pos_sorted, int64_t n, int32_t key_bytes,
                                   const double* xyzw, const double* cog3, int64_t nvoxels_host, double* min_dist,
                                   uint32_t* argmin_idx, void* ws, size_t ws_bytes, void* stream) {
  if (n < 0 || nvoxels_host < 0) return vapc_set_error(VAPC_E_INVALID, "vapc_closest_to_cog: bad arguments");
  if (n == 0) return VAPC_OK;
  if (!keys_sorted || !pos_sorted || !xyzw || !cog3 || !min_dist || !argmin_idx)
    return vapc_set_error(VAPC_E_INVALID, "vapc_closest_to_cog: null buffer");
  if (!ws || ws_bytes < seg_ws_bytes(n)) return vapc_set_error(VAPC_E_WORKSPACE, "vapc_closest_to_cog: workspace too small");
  const SegWs w = carve(ws, n);
  const int64_t nt = seg_tiles(n);
  cudaStream_t s = as_stream(stream);
  const double4* rec = reinterpret_cast<const double4*>(xyzw);
  if (key_bytes == 4) {
    typedef uint32_t K;
    const size_t smem = tile_smem_bytes<K>();
    VAPC_RETURN_IF_CUDA(cudaFuncSetAttribute(closest_kernel<K>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    VAPC_LAUNCH(closest_kernel<K>, (unsigned)nt, kThreads, smem, s, static_cast<const K*>(keys_sorted), pos_sorted, n, rec, cog3, w.tile_heads,
                                                          nvoxels_host, min_dist, argmin_idx, w.carry_cnt, w.carry);
  } else if (key_bytes == 8) {
    typedef unsigned long long K;
    const size_t smem = tile_smem_bytes<K>();
    VAPC_RETURN_IF_CUDA(cudaFuncSetAttribute(closest_kernel<K>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    VAPC_LAUNCH(closest_kernel<K>, (unsigned)nt, kThreads, smem, s, static_cast<const K*>(keys_sorted), pos_sorted, n, rec, cog3, w.tile_heads,
                                                          nvoxels_host, min_dist, argmin_idx, w.carry_cnt, w.carry);
  } else {
    return vapc_set_error(VAPC_E_INVALID, "vapc_closest_to_cog: key_bytes must be 4 or 8");
  }
  VAPC_LAUNCH(fixup_closest_kernel, (unsigned)((nt + kThreads - 1) / kThreads), kThreads, 0, s, w.tile_heads, w.carry_cnt, w.carry, (int)nt,
                                                                                     min_dist, argmin_idx);
  VAPC_CHECK_LAUNCH();
  return VAPC_OK;
}

extern "C" int vapc_merge_partials(const void* keys_sorted, const uint32_t* order, int64_t nrows, int32_t key_bytes,
                                   const uint32_t* count_in, const double* mean3_in, const double* m2_in, int64_t in_stride,
                                   int64_t nvoxels_host, void* voxel_keys, uint32_t* count, double* mean3, double* m2, void* ws,
                                   size_t ws_bytes, void* stream) {
  if (nrows < 0 || nvoxels_host < 0) return vapc_set_error(VAPC_E_INVALID, "vapc_merge_partials: bad arguments");
  if (nrows == 0) return VAPC_OK;
  if (!keys_sorted || !order || !count_in || !mean3_in || !m2_in || !voxel_keys || !count || !mean3 || !m2)
    return vapc_set_error(VAPC_E_INVALID, "vapc_merge_partials: null buffer");
  if (!ws || ws_bytes < seg_ws_bytes(nrows)) return vapc_set_error(VAPC_E_WORKSPACE, "vapc_merge_partials: workspace too small");
  const SegWs w = carve(ws, nrows);
  const int64_t nt = seg_tiles(nrows);
  cudaStream_t s = as_stream(stream);
  if (key_bytes == 4)
    VAPC_LAUNCH((merge_partials_kernel<uint32_t, kSrcSoa>), (unsigned)nt, kThreads, 0, s, static_cast<const uint32_t*>(keys_sorted), order, nrows, count_in,
        mean3_in, m2_in, in_stride, nullptr, 0, 0, w.tile_heads, nvoxels_host, static_cast<uint32_t*>(voxel_keys), count, mean3, m2);
  else if (key_bytes == 8)
    VAPC_LAUNCH((merge_partials_kernel<unsigned long long, kSrcSoa>), (unsigned)nt, kThreads, 0, s, static_cast<const unsigned long long*>(keys_sorted), order,
        nrows, count_in, mean3_in, m2_in, in_stride, nullptr, 0, 0, w.tile_heads, nvoxels_host, static_cast<unsigned long long*>(voxel_keys), count,
        mean3, m2);
  else
    return vapc_set_error(VAPC_E_INVALID, "vapc_merge_partials: key_bytes must be 4 or 8");
  VAPC_CHECK_LAUNCH();
  return VAPC_OK;
}

extern "C" int vapc_global_keys(const void* voxel_keys, int64_t nvoxels, const vapc_grid_t* local_grid_host, const vapc_grid_t* global_grid_host,
                               int32_t hist_bits, uint64_t* keys_out, uint32_t* hist, void* stream) {
  if (nvoxels < 0 || !local_grid_host || !global_grid_host || hist_bits < 1 || hist_bits > 24 || !hist)
    return vapc_set_error(VAPC_E_INVALID, "vapc_global_keys: bad arguments");
  cudaStream_t s = as_stream(stream);
  VAPC_RETURN_IF_CUDA(cudaMemsetAsync(hist, 0, sizeof(uint32_t) << hist_bits, s));
  if (nvoxels == 0) return VAPC_OK;
  if (!voxel_keys || !keys_out) return vapc_set_error(VAPC_E_INVALID, "vapc_global_keys: null buffer");
  const Grid lg = to_device_grid(local_grid_host), gg = to_device_grid(global_grid_host);
  const int global_bits = gg.bx + gg.by + gg.bz;
  if (global_bits > 64) return vapc_set_error(VAPC_E_KEYSPACE, "vapc_global_keys: global layout needs more than 64 bits");
  const int shift = global_bits > hist_bits ? global_bits - hist_bits : 0;
  const int grid = (int)std::min<int64_t>((nvoxels + kThreads - 1) / kThreads, (int64_t)kNumSMs * 16);
  unsigned long long* ko = reinterpret_cast<unsigned long long*>(keys_out);
  if (local_grid_host->key_bytes == 4)
    VAPC_LAUNCH(global_keys_kernel<uint32_t>, grid, kThreads, 0, s, static_cast<const uint32_t*>(voxel_keys), nvoxels, lg, gg, shift, ko, hist);
  else if (local_grid_host->key_bytes == 8)
    VAPC_LAUNCH(global_keys_kernel<unsigned long long>, grid, kThreads, 0, s, static_cast<const unsigned long long*>(voxel_keys), nvoxels, lg, gg, shift, ko, hist);
  else
    return vapc_set_error(VAPC_E_INVALID, "vapc_global_keys: key_bytes must be 4 or 8");
  VAPC_CHECK_LAUNCH();
  return VAPC_OK;
}

extern "C" int vapc_pack_partial_rows(const uint64_t* global_keys, const uint32_t* count, const double* mean3, const double* m2, int64_t nvoxels,
                                      int64_t begin, int64_t end, double* rows_out, void* stream) {
  if (nvoxels < 0 || begin < 0 || end > nvoxels || begin > end) return vapc_set_error(VAPC_E_INVALID, "vapc_pack_partial_rows: bad range");
  if (begin == end) return VAPC_OK;
  if (!global_keys || !count || !mean3 || !m2 || !rows_out) return vapc_set_error(VAPC_E_INVALID, "vapc_pack_partial_rows: null buffer");
  const int grid = (int)std::min<int64_t>((end - begin + kThreads - 1) / kThreads, (int64_t)kNumSMs * 16);
  VAPC_LAUNCH(pack_partial_rows_kernel, grid, kThreads, 0, as_stream(stream), reinterpret_cast<const unsigned long long*>(global_keys), count, mean3, m2,
              nvoxels, begin, end, rows_out);
  VAPC_CHECK_LAUNCH();
  return VAPC_OK;
}

extern "C" int vapc_merge_partial_rows(const void* keys_sorted, const uint32_t* order, int64_t nrows, int32_t key_bytes, const double* rows,
                                       const uint32_t* local_count, const double* local_mean3, const double* local_m2, int64_t local_stride,
                                       int64_t local_first, int64_t n_local, int64_t nvoxels_host, void* voxel_keys, uint32_t* count,
                                       double* mean3, double* m2, void* ws, size_t ws_bytes, void* stream) {
  if (nrows < 0 || nvoxels_host < 0 || n_local < 0 || n_local > nrows) return vapc_set_error(VAPC_E_INVALID, "vapc_merge_partial_rows: bad arguments");
  if (nrows == 0) return VAPC_OK;
  if (!keys_sorted || !order || !voxel_keys || !count || !mean3 || !m2 || (nrows > n_local && !rows) ||
      (n_local > 0 && (!local_count || !local_mean3 || !local_m2)))
    return vapc_set_error(VAPC_E_INVALID, "vapc_merge_partial_rows: null buffer");
  if (!ws || ws_bytes < seg_ws_bytes(nrows)) return vapc_set_error(VAPC_E_WORKSPACE, "vapc_merge_partial_rows: workspace too small");
  const SegWs w = carve(ws, nrows);
  const int64_t nt = seg_tiles(nrows);
  cudaStream_t s = as_stream(stream);
  if (key_bytes == 4)
    VAPC_LAUNCH((merge_partials_kernel<uint32_t, kSrcMixed>), (unsigned)nt, kThreads, 0, s, static_cast<const uint32_t*>(keys_sorted), order, nrows,
                local_count, local_mean3, local_m2, local_stride, rows, n_local, local_first, w.tile_heads, nvoxels_host,
                static_cast<uint32_t*>(voxel_keys), count, mean3, m2);
  else if (key_bytes == 8)
    VAPC_LAUNCH((merge_partials_kernel<unsigned long long, kSrcMixed>), (unsigned)nt, kThreads, 0, s, static_cast<const unsigned long long*>(keys_sorted),
                order, nrows, local_count, local_mean3, local_m2, local_stride, rows, n_local, local_first, w.tile_heads, nvoxels_host,
                static_cast<unsigned long long*>(voxel_keys), count, mean3, m2);
  else
    return vapc_set_error(VAPC_E_INVALID, "vapc_merge_partial_rows: key_bytes must be 4 or 8");
  VAPC_CHECK_LAUNCH();
  return VAPC_OK;
}

extern "C" int vapc_merge_sorted_tables_stats(const uint64_t* own_keys, const uint32_t* own_count, const double* own_mean3, const double* own_m2,
                                              int64_t own_stride, int64_t own_first, int64_t n_own, const uint64_t* u_keys, const uint32_t* u_count,
                                              const double* u_mean3, const double* u_m2, int64_t nu, const int64_t* u_new_before, int32_t key_bytes,
                                              int64_t nvoxels_host, void* out_keys, uint32_t* out_count, double* out_mean3, double* out_m2,
                                              const vapc_grid_t* grid_host, const int32_t* axis_order_host, double* density, double* cog3,
                                              double* std3, double* cov6, double* eig3, double* eign3, double* feat8, void* stream) {
  if (n_own < 0 || nu < 0 || nvoxels_host < n_own || own_first < 0 || own_stride < own_first + n_own)
    return vapc_set_error(VAPC_E_INVALID, "vapc_merge_sorted_tables: bad arguments");
  if (key_bytes != 4 && key_bytes != 8) return vapc_set_error(VAPC_E_INVALID, "vapc_merge_sorted_tables: key_bytes must be 4 or 8");
  StatsOut so{density, cog3, std3, cov6, eig3, eign3, feat8, nvoxels_host, {}};
  const int with_stats = density || cog3 || so.needs_m2();
  if (with_stats && !grid_host) return vapc_set_error(VAPC_E_INVALID, "vapc_merge_sorted_tables_stats: statistics need the grid");
  if (!stats_axes_from_order(axis_order_host, &so.ax)) return vapc_set_error(VAPC_E_INVALID, "vapc_merge_sorted_tables_stats: axis_order must be a permutation of 0, 1, 2");
  if (nvoxels_host == 0) return VAPC_OK;
  if (!out_keys || !out_count || (n_own > 0 && (!own_keys || !own_count || !own_mean3 || !own_m2)) ||
      (nu > 0 && (!u_keys || !u_count || !u_mean3 || !u_m2 || !u_new_before)))
    return vapc_set_error(VAPC_E_INVALID, "vapc_merge_sorted_tables: null buffer");
  Grid g{};
  if (grid_host) g = to_device_grid(grid_host);
  cudaStream_t s = as_stream(stream);
  typedef unsigned long long u64;
  const u64* ok = reinterpret_cast<const u64*>(own_keys);
  const u64* uk = reinterpret_cast<const u64*>(u_keys);
  const unsigned g1 = (unsigned)((n_own + kThreads * kMergeChunks - 1) / (kThreads * kMergeChunks)), g2 = (unsigned)((nu + kThreads - 1) / kThreads);
  if (key_bytes == 4) {
    if (n_own) VAPC_LAUNCH(merge_own_rows_kernel<uint32_t>, g1, kThreads, 0, s, ok, own_count, own_mean3, own_m2, own_stride, own_first, n_own, uk, u_count, u_mean3, u_m2,
                           nu, u_new_before, nvoxels_host, static_cast<uint32_t*>(out_keys), out_count, out_mean3, out_m2, g, so, with_stats);
    if (nu) VAPC_LAUNCH(merge_u_rows_kernel<uint32_t>, g2, kThreads, 0, s, ok, own_count, own_mean3, own_m2, own_stride, own_first, n_own, uk, u_count, u_mean3, u_m2, nu, u_new_before, nvoxels_host,
                        static_cast<uint32_t*>(out_keys), out_count, out_mean3, out_m2, g, so, with_stats);
  } else {
    if (n_own) VAPC_LAUNCH(merge_own_rows_kernel<u64>, g1, kThreads, 0, s, ok, own_count, own_mean3, own_m2, own_stride, own_first, n_own, uk, u_count, u_mean3, u_m2, nu,
                           u_new_before, nvoxels_host, static_cast<u64*>(out_keys), out_count, out_mean3, out_m2, g, so, with_stats);
    if (nu) VAPC_LAUNCH(merge_u_rows_kernel<u64>, g2, kThreads, 0, s, ok, own_count, own_mean3, own_m2, own_stride, own_first, n_own, uk, u_count, u_mean3, u_m2, nu, u_new_before, nvoxels_host,
                        static_cast<u64*>(out_keys), out_count, out_mean3, out_m2, g, so, with_stats);
  }
  VAPC_CHECK_LAUNCH();
  return VAPC_OK;
}

extern "C" int vapc_merge_sorted_tables(const uint64_t* own_keys, const uint32_t* own_count, const double* own_mean3, const double* own_m2,
                                        int64_t own_stride, int64_t own_first, int64_t n_own, const uint64_t* u_keys, const uint32_t* u_count,
                                        const double* u_mean3, const double* u_m2, int64_t nu, const int64_t* u_new_before, int32_t key_bytes,
                                        int64_t nvoxels_host, void* out_keys, uint32_t* out_count, double* out_mean3, double* out_m2, void* stream) {
  if (nvoxels_host > 0 && (!out_mean3 || !out_m2)) return vapc_set_error(VAPC_E_INVALID, "vapc_merge_sorted_tables: null buffer");
  return vapc_merge_sorted_tables_stats(own_keys, own_count, own_mean3, own_m2, own_stride, own_first, n_own, u_keys, u_count, u_mean3, u_m2, nu, u_new_before,
                                        key_bytes, nvoxels_host, out_keys, out_count, out_mean3, out_m2, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr,
                                        nullptr, nullptr, nullptr, stream);
}
