// Counting sort of the key-prefix buckets through a per-bucket CELL TABLE in shared memory ("dense grids").
//
// After vapc_pack_records_bucketed a bucket holds the points whose key starts with the same 8 bits; what is left to sort
// are the low_bits below.  When those are <= 16 the whole key range of a bucket — 2^low_bits cells — fits a table of 16-bit
// counters in one SM's shared memory (128 KB), and the two radix passes + digit histograms + head count of the sort path
// (1.4 ms of a 6.7 ms step at 1e8 points, each bound by shared-memory ranking wavefronts at 34 % of the HBM roofline) become:
//
//   one CTA (1024 threads) per bucket
//   1 count    stream the bucket's keys, shared-memory atomic per key on its cell's half-word            4 B/pt read
//   2 scan     cells -> first slot: 64 cells (32 words) per warp step, 16-bit offsets relative to a 32-bit base per group
//   3 place    stream the keys again, atomic-with-return on the cell = the slot; the point's bucket position goes there
//                                                                                 4 B/pt read, 4 B/pt scattered inside the bucket's own window
//   4 fix      thread per cell: head counts per 1024-tile (what vapc_count_voxels would have produced), the cell's key to
//              its slots, and the canonical order: the atomics of step 3 hand out a cell's slots in arrival order, so a
//              cell whose positions are not ascending is sorted (<= 16: Batcher network in registers, <= 32: the same with
//              32, <= kMaxCellCount: ranks counted by a warp).  Ascending positions = input order inside the voxel, i.e.
//              the output is bit-identical to the stable radix sort (tests/test_gpu_kernels.py::test_cell_sort_equals_radix).
//
// Not applicable -> status bit 2 and nothing else is promised (the caller falls back to vapc_sort_pairs_segmented): a
// cell with more than kMaxCellCount points (the per-cell ordering would dominate) or a 16-bit counter overflow.
#include <algorithm>

#include "segment_ws.cuh"

namespace {
using namespace segws;
constexpr int kCsThreads = 1024;
constexpr int kCsWarps = kCsThreads / 32;
constexpr int kMaxCellBits = 16;
constexpr int kMinCellBits = 7;      // one group of 64 cells at least
constexpr int kMaxCellCount = 128;   // beyond that a voxel's ordering is the radix sort's job
constexpr int kFixBlock = 4096;      // cells per round of step 4 (worklists are sized for it)
constexpr int kLoadUnroll = 8;

struct CsSmem {
  unsigned gbase[1 << (kMaxCellBits - 6)];  // first slot of every group of 64 cells
  unsigned warp_tot[kCsWarps];
  unsigned wl_n[3];
  unsigned flags;
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

// first slot / number of points of cell c after step 3 (every half-word then holds the END of its cell inside the group)
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

__global__ void __launch_bounds__(kCsThreads, 1) cell_sort_kernel(const uint32_t* __restrict__ keys_b, const uint32_t* __restrict__ bucket_start,
                                                                  int low_bits, uint32_t* __restrict__ keys_sorted, uint32_t* __restrict__ pos_sorted,
                                                                  uint32_t* __restrict__ tile_heads, int* __restrict__ status) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  CsSmem& sm = *reinterpret_cast<CsSmem*>(smem_raw);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const unsigned b = blockIdx.x;
  const unsigned s0 = bucket_start[b], s1 = bucket_start[b + 1];
  if (s0 == s1) return;
  const unsigned ncells = 1u << low_bits, cmask = ncells - 1u;
  const unsigned nwords = ncells >> 1, ngroups = ncells >> 6;

  for (unsigned i = tid; i < nwords; i += kCsThreads) sm.w[i] = 0;
  if (tid == 0) sm.flags = 0;
  __syncthreads();

  // ---- 1 count ------------------------------------------------------------------------------------------------
  for (unsigned base = s0; base < s1; base += kCsThreads * kLoadUnroll) {  // s1 - s0 < 2^32 - 8192 is guaranteed by n < 2^32 - 2^13 (checked by the host)
    unsigned k[kLoadUnroll];
#pragma unroll
    for (int u = 0; u < kLoadUnroll; ++u) {
      const unsigned i = base + u * kCsThreads + tid;
      k[u] = i < s1 ? __ldcs(keys_b + i) : 0u;
    }
#pragma unroll
    for (int u = 0; u < kLoadUnroll; ++u) {
      const unsigned i = base + u * kCsThreads + tid;
      if (i < s1) {
        const unsigned c = k[u] & cmask;
        atomicAdd(&sm.w[c >> 1], 1u << ((c & 1u) << 4));
      }
    }
  }
  __syncthreads();

  // ---- 2 scan: warp w owns the groups [w * gpw, (w + 1) * gpw) ------------------------------------------------------
  const unsigned gpw = (ngroups + kCsWarps - 1) / kCsWarps;
  const unsigned g0 = min(ngroups, warp * gpw), g1 = min(ngroups, g0 + gpw);
  {
    unsigned tot = 0, mx = 0;
    for (unsigned g = g0; g < g1; ++g) {
      const unsigned word = sm.w[g * 32 + lane];
      const unsigned a = word & 0xffffu, c2 = word >> 16;
      tot += a + c2;
      mx = max(mx, max(a, c2));
    }
    tot = __reduce_add_sync(0xffffffffu, tot);
    mx = __reduce_max_sync(0xffffffffu, mx);
    if (lane == 0) {
      sm.warp_tot[warp] = tot;
      if (mx > (unsigned)kMaxCellCount) atomicOr(&sm.flags, 1u);
    }
  }
  __syncthreads();
  if (warp == 0) {
    const unsigned t = sm.warp_tot[lane];
    unsigned inc = t;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const unsigned u = __shfl_up_sync(0xffffffffu, inc, o);
      if (lane >= o) inc += u;
    }
    sm.warp_tot[lane] = inc - t;
    // a half-word that overflowed carried into its neighbour: the total no longer equals the bucket's size
    if (lane == 31 && inc != s1 - s0) atomicOr(&sm.flags, 1u);
  }
  __syncthreads();
  if (sm.flags) {
    if (tid == 0) atomicOr(status, 4);
    return;
  }
  {
    unsigned running = s0 + sm.warp_tot[warp];
    for (unsigned g = g0; g < g1; ++g) {
      const unsigned word = sm.w[g * 32 + lane];
      const unsigned a = word & 0xffffu, s = a + (word >> 16);
      unsigned inc = s;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const unsigned u = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += u;
      }
      const unsigned e = inc - s;
      sm.w[g * 32 + lane] = e | ((e + a) << 16);  // group totals are <= 64 * kMaxCellCount: 16 bits hold every offset
      if (lane == 0) sm.gbase[g] = running;
      running += __shfl_sync(0xffffffffu, inc, 31);
    }
  }
  __syncthreads();

  // ---- 3 place ------------------------------------------------------------------------------------------------
  for (unsigned base = s0; base < s1; base += kCsThreads * kLoadUnroll) {
    unsigned k[kLoadUnroll];
#pragma unroll
    for (int u = 0; u < kLoadUnroll; ++u) {
      const unsigned i = base + u * kCsThreads + tid;
      k[u] = i < s1 ? __ldcs(keys_b + i) : 0u;
    }
#pragma unroll
    for (int u = 0; u < kLoadUnroll; ++u) {
      const unsigned i = base + u * kCsThreads + tid;
      if (i < s1) {
        const unsigned c = k[u] & cmask;
        const unsigned sh = (c & 1u) << 4;
        const unsigned old = atomicAdd(&sm.w[c >> 1], 1u << sh);
        pos_sorted[sm.gbase[c >> 6] + ((old >> sh) & 0xffffu)] = i;
      }
    }
  }
  __syncthreads();

  // ---- 4 heads, keys, canonical order -----------------------------------------------------------------------------
  const unsigned key_high = b << low_bits;
  for (unsigned blk = 0; blk < ncells; blk += kFixBlock) {
    if (tid < 3) sm.wl_n[tid] = 0;
    __syncthreads();
    const unsigned blk_cells = min((unsigned)kFixBlock, ncells - blk);
    for (unsigned j = 0; j < blk_cells; j += kCsThreads) {  // warp-uniform trip count: blk_cells is a multiple of 64 ... of 1024 when >= 1024
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
      unsigned v[kPer], rank[kPer];
#pragma unroll
      for (int r = 0; r < kPer; ++r) {
        v[r] = r * 32 + lane < n ? __ldcg(pos_sorted + start + r * 32 + lane) : 0xffffffffu;
        rank[r] = 0;
      }
#pragma unroll
      for (int r2 = 0; r2 < kPer; ++r2) {
        for (int q = 0; q < 32; ++q) {
          const unsigned u = __shfl_sync(0xffffffffu, v[r2], q);  // positions are distinct; the padding never counts as smaller
#pragma unroll
          for (int r = 0; r < kPer; ++r) rank[r] += u < v[r] ? 1u : 0u;
        }
      }
      __syncwarp();
#pragma unroll
      for (int r = 0; r < kPer; ++r)
        if (r * 32 + lane < n) pos_sorted[start + rank[r]] = v[r];
    }
    __syncthreads();
  }
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
  VAPC_LAUNCH(cell_sort_kernel, (unsigned)nseg, kCsThreads, smem, s, static_cast<const uint32_t*>(keys_bucketed), bucket_start, low_bits,
              static_cast<uint32_t*>(keys_sorted), pos_sorted, w.tile_heads, status);
  VAPC_RETURN_IF_CUDA(scan_u32_exclusive(w.tile_heads, w.tile_heads, nt, w.total, w.scan_ws, s));
  VAPC_RETURN_IF_CUDA(cudaMemcpyAsync(nvoxels, w.total, sizeof(int64_t), cudaMemcpyDeviceToDevice, s));
  VAPC_CHECK_LAUNCH();
  return VAPC_OK;
}
