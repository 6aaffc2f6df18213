// Counting sort of the key-prefix buckets through a per-bucket CELL TABLE in (distributed) shared memory — "dense grids".
//
// After vapc_pack_records_bucketed a bucket holds the points whose key starts with the same 8 bits; what is left to sort are
// the low_bits below.  When those are <= 16 the whole key range of a bucket — 2^low_bits cells — fits a table of 16-bit counters
// in one SM's shared memory (128 KB), and the two radix passes + digit histograms + head count of the sort path (1.4 ms of a
// 6.7 ms step at 1e8 points, each bound by shared-memory ranking wavefronts at 34 % of the HBM roofline) become one kernel.
//
// One CLUSTER of 8 CTAs (1024 threads each, one per SM) per bucket; CTA r streams the r-th eighth of the bucket's keys:
//   1 count    shared-memory atomic per key on its cell's half-word in the CTA's OWN table                       4 B/pt read
//   2 scan     CTA r owns the r-th slice of the cells: it reads that slice of all 8 tables through distributed shared memory,
//              scans it, and writes back to every CTA the first slot of ITS points of every cell (cell start + the counts of the
//              CTAs in front: CTA r holds earlier input positions than CTA r + 1) as 16-bit offsets relative to a 32-bit base
//              per group of 64 cells
//   3 place    stream the keys again, atomic-with-return on the cell of the own table = the slot; the point's bucket position
//              goes there: a 4-byte scatter inside the bucket's own 1.5 MB window of pos_sorted.  (Why a cluster: with one CTA
//              per bucket 148 such windows are open at once — 220 MB, twice the L2 — and the scatter ran at DRAM sector
//              granularity, 6.6 GB of DRAM traffic for 1.6 GB of work, 4.1 ms, profiles/r2_a.  16 clusters keep 24 MB open.)
//   4 fix      CTA r, thread per cell of its slice: head counts per 1024-tile (what vapc_count_voxels would have produced), the
//              cell's key to its slots, and the canonical order: the atomics of step 3 hand out a cell's slots in arrival order,
//              so a cell whose positions are not ascending is sorted (<= 16: Batcher network in registers, <= 32: the same with
//              32, <= kMaxCellCount: ranks counted by a warp).  Ascending positions = input order inside the voxel, i.e. the
//              output is bit-identical to the stable radix sort (tests/test_gpu_kernels.py::test_cell_sort_equals_radix).
//
// Not applicable -> status bit 2 and nothing else is promised (the caller falls back to vapc_sort_pairs_segmented): a
// cell with more than kMaxCellCount points (the per-cell ordering would dominate) or a 16-bit counter overflow.
#include <cooperative_groups.h>

#include <algorithm>

#include "segment_ws.cuh"

namespace {
namespace cg = cooperative_groups;
using namespace segws;
constexpr int kCsThreads = 1024;
constexpr int kCsWarps = kCsThreads / 32;
constexpr int kRanks = 8;            // CTAs per cluster (the portable maximum)
constexpr int kMaxCellBits = 16;
constexpr int kMinCellBits = 9;      // one group of 64 cells per CTA of the cluster at least
constexpr int kMaxCellCount = 128;   // beyond that a voxel's ordering is the radix sort's job
constexpr int kFixBlock = 4096;      // cells per round of step 4 (worklists are sized for it)
constexpr int kLoadUnroll = 8;
constexpr int kMaxGroups = 1 << (kMaxCellBits - 6);

struct CsSmem {
  unsigned gbase[kMaxGroups];        // first slot of every group of 64 cells (all groups: the CTA's keys fall anywhere)
  unsigned gtot[kMaxGroups / kRanks];  // points of every group of the CTA's own slice
  unsigned slice_tot[kRanks];        // points of every slice (written by its owner into every CTA)
  unsigned slice_flag[kRanks];       // "not applicable" raised by the owner of a slice
  unsigned warp_tot[kCsWarps], warp_flag[kCsWarps];
  unsigned wl_n[3];
  unsigned short wl16[kFixBlock], wl32[kFixBlock], wlbig[kFixBlock];
  alignas(16) unsigned w[1];  // 2^low_bits / 2 words: two 16-bit cells per word (dynamic)
};
inline size_t cs_smem_bytes(int low_bits) { return sizeof(CsSmem) + (((size_t)1 << low_bits) / 2) * sizeof(unsigned); }

__device__ __forceinline__ void ce(unsigned& a, unsigned& b) {
  const unsigned lo = min(a, b), hi = max(a, b);
  a = lo;
  b = hi;
}
// Batcher's odd-even merge sort as a fully unrolled compare-exchange network (N a power of two)
template <int N>
__device__ __forceinline__ void network_sort(unsigned (&v)[N]) {
#pragma unroll
  for (int p = 1; p < N; p <<= 1) {
#pragma unroll
    for (int k = p; k >= 1; k >>= 1) {
#pragma unroll
      for (int j = k % p; j + k < N; j += 2 * k) {
#pragma unroll
        for (int i = 0; i < k; ++i) {
          if (i + j + k < N && (i + j) / (2 * p) == (i + j + k) / (2 * p)) ce(v[i + j], v[i + j + k]);
        }
      }
    }
  }
}

template <int N>
__device__ __forceinline__ void sort_cell_in_registers(uint32_t* __restrict__ pos, unsigned start, int n) {
  unsigned v[N];
#pragma unroll
  for (int j = 0; j < N; ++j) v[j] = j < n ? __ldcg(pos + start + j) : 0xffffffffu;
  network_sort<N>(v);
#pragma unroll
  for (int j = 0; j < N; ++j)
    if (j < n) pos[start + j] = v[j];
}

// first slot / number of points of cell c once the table holds, per half-word, the END of the cell inside its group
__device__ __forceinline__ void cell_range(const CsSmem& sm, unsigned c, unsigned& start, int& n) {
  const unsigned word = sm.w[c >> 1];
  unsigned end_rel, start_rel;
  if (c & 1u) {
    end_rel = word >> 16;
    start_rel = word & 0xffffu;
  } else {
    end_rel = word & 0xffffu;
    start_rel = (c & 63u) == 0 ? 0u : (sm.w[(c >> 1) - 1] >> 16);
  }
  start = sm.gbase[c >> 6] + start_rel;
  n = (int)(end_rel - start_rel);
}

__global__ void __cluster_dims__(kRanks, 1, 1) __launch_bounds__(kCsThreads, 1)
    cell_sort_kernel(const uint32_t* __restrict__ keys_b, const uint32_t* __restrict__ bucket_start, int low_bits, uint32_t* __restrict__ keys_sorted,
                     uint32_t* __restrict__ pos_sorted, uint32_t* __restrict__ tile_heads, int* __restrict__ status) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  CsSmem& sm = *reinterpret_cast<CsSmem*>(smem_raw);
  cg::cluster_group cluster = cg::this_cluster();
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const unsigned rank = cluster.block_rank();
  const unsigned b = blockIdx.x / kRanks;
  const unsigned s0 = bucket_start[b], s1 = bucket_start[b + 1];
  if (s0 == s1) return;  // the whole cluster leaves together
  const unsigned ncells = 1u << low_bits, cmask = ncells - 1u;
  const unsigned nwords = ncells >> 1, ngroups = ncells >> 6;
  // this CTA's share of the bucket's keys (input order: rank 0 first), and of the cells
  const unsigned share = (s1 - s0 + kRanks - 1) / kRanks;
  const unsigned k0 = min(s1, s0 + rank * share), k1 = min(s1, k0 + share);
  const unsigned gslice = ngroups / kRanks;  // groups per slice (ngroups is a power of two >= kRanks)
  const unsigned gs0 = rank * gslice;

  for (unsigned i = tid; i < nwords; i += kCsThreads) sm.w[i] = 0;
  __syncthreads();

  // ---- 1 count ------------------------------------------------------------------------------------------------
  for (unsigned base = k0; base < k1; base += kCsThreads * kLoadUnroll) {  // no wrap: n < 2^32 - 2^13 (checked by the host)
    unsigned k[kLoadUnroll];
#pragma unroll
    for (int u = 0; u < kLoadUnroll; ++u) {
      const unsigned i = base + u * kCsThreads + tid;
      k[u] = i < k1 ? __ldcs(keys_b + i) : 0u;
    }
#pragma unroll
    for (int u = 0; u < kLoadUnroll; ++u) {
      const unsigned i = base + u * kCsThreads + tid;
      if (i < k1) {
        const unsigned c = k[u] & cmask;
        atomicAdd(&sm.w[c >> 1], 1u << ((c & 1u) << 4));
      }
    }
  }
  cluster.sync();

  // ---- 2 scan of the own slice over all 8 tables -------------------------------------------------------------------------
  unsigned* tab[kRanks];
#pragma unroll
  for (int q = 0; q < kRanks; ++q) tab[q] = cluster.map_shared_rank(sm.w, q);
  // 2a: points per group of the slice (a warp per group, a lane per word), largest cell, overflow check
  {
    unsigned mx = 0, tot = 0;
    for (unsigned g = warp; g < gslice; g += kCsWarps) {
      const unsigned wi = (gs0 + g) * 32 + lane;
      unsigned a = 0, c2 = 0;
#pragma unroll
      for (int q = 0; q < kRanks; ++q) {
        const unsigned word = tab[q][wi];
        a += word & 0xffffu;
        c2 += word >> 16;
      }
      mx = max(mx, max(a, c2));
      const unsigned s = __reduce_add_sync(0xffffffffu, a + c2);
      if (lane == 0) sm.gtot[g] = s;
      tot += s;  // the same in every lane
    }
    mx = __reduce_max_sync(0xffffffffu, mx);
    if (lane == 0) {
      sm.warp_tot[warp] = tot;
      sm.warp_flag[warp] = mx > (unsigned)kMaxCellCount ? 1u : 0u;
    }
  }
  __syncthreads();
  if (warp == 0) {
    const unsigned tot = __reduce_add_sync(0xffffffffu, sm.warp_tot[lane]);
    const unsigned flag = __reduce_or_sync(0xffffffffu, sm.warp_flag[lane]);
    // slice total and flag into every CTA of the cluster
    if (lane < kRanks) {
      cluster.map_shared_rank(sm.slice_tot, lane)[rank] = tot;
      cluster.map_shared_rank(sm.slice_flag, lane)[rank] = flag;
    }
  }
  // exclusive prefix of the group totals inside the slice (<= 128 groups: warp 1, four per lane)
  if (warp == 1) {
    unsigned run = 0;
    for (unsigned g0 = 0; g0 < gslice; g0 += 32) {
      const unsigned g = g0 + lane;
      const unsigned t = g < gslice ? sm.gtot[g] : 0u;
      unsigned inc = t;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const unsigned u = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += u;
      }
      if (g < gslice) sm.gtot[g] = run + inc - t;
      run += __shfl_sync(0xffffffffu, inc, 31);
    }
  }
  __syncthreads();
  // 2b: per cell of the slice the first slot of every CTA's points, 16 bits relative to the group; the group's slice-relative
  // base to every CTA's gbase
  for (unsigned g = warp; g < gslice; g += kCsWarps) {
    const unsigned wi = (gs0 + g) * 32 + lane;
    unsigned wa[kRanks], wb[kRanks];
    unsigned a = 0, c2 = 0;
#pragma unroll
    for (int q = 0; q < kRanks; ++q) {
      const unsigned word = tab[q][wi];
      wa[q] = word & 0xffffu;
      wb[q] = word >> 16;
      a += wa[q];
      c2 += wb[q];
    }
    const unsigned s = a + c2;
    unsigned inc = s;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const unsigned u = __shfl_up_sync(0xffffffffu, inc, o);
      if (lane >= o) inc += u;
    }
    unsigned ea = inc - s, eb = ea + a;  // first slot of the even / odd cell inside the group
#pragma unroll
    for (int q = 0; q < kRanks; ++q) {
      tab[q][wi] = ea | (eb << 16);  // group totals are <= 64 * kMaxCellCount (else the flag is up): 16 bits hold every offset
      ea += wa[q];
      eb += wb[q];
    }
    if (lane < kRanks) cluster.map_shared_rank(sm.gbase, lane)[gs0 + g] = sm.gtot[g];
  }
  cluster.sync();
  {
    unsigned flag = 0, total = 0;
#pragma unroll
    for (int q = 0; q < kRanks; ++q) {
      flag |= sm.slice_flag[q];
      total += sm.slice_tot[q];
    }
    // a half-word that overflowed carried into its neighbour: the total no longer equals the bucket's size
    if (flag || total != s1 - s0) {
      if (tid == 0 && rank == 0) atomicOr(status, 4);
      cluster.sync();  // nobody leaves while its table may still be read
      return;
    }
  }
  // group bases: slice-relative -> absolute
  for (unsigned g = tid; g < ngroups; g += kCsThreads) {
    const unsigned q = g / gslice;
    unsigned base = s0;
    for (unsigned q2 = 0; q2 < q; ++q2) base += sm.slice_tot[q2];
    sm.gbase[g] += base;
  }
  __syncthreads();

  // ---- 3 place ------------------------------------------------------------------------------------------------
  for (unsigned base = k0; base < k1; base += kCsThreads * kLoadUnroll) {
    unsigned k[kLoadUnroll];
#pragma unroll
    for (int u = 0; u < kLoadUnroll; ++u) {
      const unsigned i = base + u * kCsThreads + tid;
      k[u] = i < k1 ? __ldcs(keys_b + i) : 0u;
    }
#pragma unroll
    for (int u = 0; u < kLoadUnroll; ++u) {
      const unsigned i = base + u * kCsThreads + tid;
      if (i < k1) {
        const unsigned c = k[u] & cmask;
        const unsigned sh = (c & 1u) << 4;
        const unsigned old = atomicAdd(&sm.w[c >> 1], 1u << sh);
        __stcg(pos_sorted + sm.gbase[c >> 6] + ((old >> sh) & 0xffffu), i);
      }
    }
  }
  __threadfence();
  cluster.sync();

  // ---- 4 heads, keys, canonical order: the cells of the own slice ---------------------------------------------------------------
  // the LAST CTA's table now holds the end of every cell (its points come last): take the slice from there
  if (rank != kRanks - 1) {
    const unsigned* last = tab[kRanks - 1];
    for (unsigned i = gs0 * 32 + tid; i < (gs0 + gslice) * 32; i += kCsThreads) sm.w[i] = last[i];
  }
  __syncthreads();
  const unsigned key_high = b << low_bits;
  const unsigned c_begin = gs0 * 64, c_end = (gs0 + gslice) * 64;
  for (unsigned blk = c_begin; blk < c_end; blk += kFixBlock) {
    if (tid < 3) sm.wl_n[tid] = 0;
    __syncthreads();
    const unsigned blk_cells = min((unsigned)kFixBlock, c_end - blk);
    for (unsigned j = 0; j < blk_cells; j += kCsThreads) {  // warp-uniform trip count
      const unsigned lc = j + tid;
      const bool in_range = lc < blk_cells;
      const unsigned c = blk + lc;
      unsigned start = 0;
      int n = 0;
      if (in_range) cell_range(sm, c, start, n);
      // head count of the tile the cell starts in: one atomic per run of lanes that share a tile (starts ascend with the lane)
      const unsigned occ = __ballot_sync(0xffffffffu, n > 0);
      const unsigned tile = start / kSegTile;
      const unsigned lower = occ & ((1u << lane) - 1u);
      const int p = lower ? 31 - __clz(lower) : -1;
      const unsigned ptile = __shfl_sync(0xffffffffu, tile, p < 0 ? 0 : p);
      const bool leader = n > 0 && (p < 0 || ptile != tile);
      const unsigned leaders = __ballot_sync(0xffffffffu, leader);
      if (leader) {
        const unsigned after = leaders & ~((2u << lane) - 1u);
        const unsigned upto = after ? ((1u << (__ffs(after) - 1)) - 1u) : 0xffffffffu;
        atomicAdd(tile_heads + tile, (unsigned)__popc(occ & upto & ~((1u << lane) - 1u)));
      }
      if (n > 0) {
        const unsigned key = key_high | c;
        for (int q = 0; q < n; ++q) keys_sorted[start + q] = key;
        if (n > 1) {
          unsigned prev = __ldcg(pos_sorted + start);
          bool bad = false;
          for (int q = 1; q < n; ++q) {
            const unsigned v = __ldcg(pos_sorted + start + q);
            bad |= v < prev;
            prev = v;
          }
          if (bad) {
            if (n <= 16) sm.wl16[atomicAdd(&sm.wl_n[0], 1u)] = (unsigned short)lc;
            else if (n <= 32) sm.wl32[atomicAdd(&sm.wl_n[1], 1u)] = (unsigned short)lc;
            else sm.wlbig[atomicAdd(&sm.wl_n[2], 1u)] = (unsigned short)lc;
          }
        }
      }
    }
    __syncthreads();
    for (unsigned i = tid; i < sm.wl_n[0]; i += kCsThreads) {
      unsigned start;
      int n;
      cell_range(sm, blk + sm.wl16[i], start, n);
      sort_cell_in_registers<16>(pos_sorted, start, n);
    }
    for (unsigned i = tid; i < sm.wl_n[1]; i += kCsThreads) {
      unsigned start;
      int n;
      cell_range(sm, blk + sm.wl32[i], start, n);
      sort_cell_in_registers<32>(pos_sorted, start, n);
    }
    // 32 < n <= kMaxCellCount: a warp per cell counts, for each of its (up to 4 per lane) values, the smaller ones
    for (unsigned i = warp; i < sm.wl_n[2]; i += kCsWarps) {
      unsigned start;
      int n;
      cell_range(sm, blk + sm.wlbig[i], start, n);
      constexpr int kPer = kMaxCellCount / 32;
      unsigned v[kPer], rnk[kPer];
#pragma unroll
      for (int r = 0; r < kPer; ++r) {
        v[r] = r * 32 + lane < n ? __ldcg(pos_sorted + start + r * 32 + lane) : 0xffffffffu;
        rnk[r] = 0;
      }
#pragma unroll
      for (int r2 = 0; r2 < kPer; ++r2) {
        for (int q = 0; q < 32; ++q) {
          const unsigned u = __shfl_sync(0xffffffffu, v[r2], q);  // positions are distinct; the padding never counts as smaller
#pragma unroll
          for (int r = 0; r < kPer; ++r) rnk[r] += u < v[r] ? 1u : 0u;
        }
      }
      __syncwarp();
#pragma unroll
      for (int r = 0; r < kPer; ++r)
        if (r * 32 + lane < n) pos_sorted[start + rnk[r]] = v[r];
    }
    __syncthreads();
  }
  cluster.sync();  // the last CTA's table was read by the others
}
}  // namespace

extern "C" int vapc_sort_cells_segmented(const void* keys_bucketed, int64_t n, int32_t key_bytes, int32_t low_bits, const uint32_t* bucket_start,
                                         int32_t nseg, void* keys_sorted, uint32_t* pos_sorted, int64_t* nvoxels, int32_t* status, void* ws,
                                         size_t ws_bytes, void* stream) {
  if (n < 0 || !nvoxels || !status) return vapc_set_error(VAPC_E_INVALID, "vapc_sort_cells_segmented: bad arguments");
  if (key_bytes != 4) return vapc_set_error(VAPC_E_INVALID, "vapc_sort_cells_segmented: 32-bit keys only");
  if (low_bits < kMinCellBits || low_bits > kMaxCellBits)
    return vapc_set_error(VAPC_E_INVALID, "vapc_sort_cells_segmented: %d..%d key bits below the bucket digit", kMinCellBits, kMaxCellBits);
  if (nseg < 1 || nseg > 256 || (nseg > 1 && low_bits + 8 > 32)) return vapc_set_error(VAPC_E_INVALID, "vapc_sort_cells_segmented: 1..256 segments");
  cudaStream_t s = as_stream(stream);
  if (n == 0) {
    VAPC_RETURN_IF_CUDA(cudaMemsetAsync(nvoxels, 0, sizeof(int64_t), s));
    return VAPC_OK;
  }
  if (n >= ((int64_t)1 << 32) - (kCsThreads * kLoadUnroll)) return vapc_set_error(VAPC_E_INVALID, "vapc_sort_cells_segmented: n must be < 2^32 - 2^13");
  if (!keys_bucketed || !bucket_start || !keys_sorted || !pos_sorted) return vapc_set_error(VAPC_E_INVALID, "vapc_sort_cells_segmented: null buffer");
  if (!ws || ws_bytes < seg_ws_bytes(n)) return vapc_set_error(VAPC_E_WORKSPACE, "vapc_sort_cells_segmented: workspace too small");
  const SegWs w = carve(ws, n);
  const int64_t nt = seg_tiles(n);
  VAPC_RETURN_IF_CUDA(cudaMemsetAsync(w.tile_heads, 0, (size_t)nt * sizeof(uint32_t), s));
  const size_t smem = cs_smem_bytes(low_bits);
  VAPC_RETURN_IF_CUDA(cudaFuncSetAttribute(cell_sort_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)cs_smem_bytes(kMaxCellBits)));
  VAPC_LAUNCH(cell_sort_kernel, (unsigned)nseg * kRanks, kCsThreads, smem, s, static_cast<const uint32_t*>(keys_bucketed), bucket_start, low_bits,
              static_cast<uint32_t*>(keys_sorted), pos_sorted, w.tile_heads, status);
  VAPC_CHECK_LAUNCH();
  VAPC_RETURN_IF_CUDA(scan_u32_exclusive(w.tile_heads, w.tile_heads, nt, w.total, w.scan_ws, s));
  VAPC_RETURN_IF_CUDA(cudaMemcpyAsync(nvoxels, w.total, sizeof(int64_t), cudaMemcpyDeviceToDevice, s));
  return VAPC_OK;
}
