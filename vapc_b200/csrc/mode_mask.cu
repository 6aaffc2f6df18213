// mode / mode_count of an integer attribute, mask membership by hashed key, voxel buffer expansion.
// Reference: vapc.py:407-433 (mode, mode_count), :594-598 (MultiIndex.isin), :495-541 (voxel buffer).
#include <math.h>

#include <algorithm>
#include <vector>

#include "np_sum.cuh"
#include "common.cuh"

namespace {
constexpr int kThreads = 256;
constexpr unsigned long long kEmpty = ~0ull;

inline int grid_for(int64_t n, int per_thread = 4) {
  int64_t b = (n + (int64_t)kThreads * per_thread - 1) / ((int64_t)kThreads * per_thread);
  const int64_t cap = (int64_t)kNumSMs * 16;
  if (b > cap) b = cap;
  return (int)(b < 1 ? 1 : b);
}

// ---- composite (voxel row, attribute value) keys ----------------------------------------------------
// astype(int) truncates toward zero (vapc.py:409, :423); negative or too-large values are flagged
// (np.bincount raises on negatives).
template <typename KeyT>
__global__ void __launch_bounds__(kThreads) mode_keys_kernel(const uint32_t* __restrict__ p2v, const double* __restrict__ attr, int64_t n,
                                                             int attr_bits, KeyT* __restrict__ keys, int* __restrict__ status) {
  int bad = 0;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    const double a = attr[i];
    const long long v = (long long)a;  // trunc toward zero
    if (!(a == a) || v < 0) bad |= 2;
    else if ((unsigned long long)v >> attr_bits) bad |= 4;
    keys[i] = (KeyT)(((unsigned long long)p2v[i] << attr_bits) | ((unsigned long long)v & ((1ull << attr_bits) - 1ull)));
  }
  if (bad) atomicOr(status, bad);
}

// One thread per voxel walks the voxel's (value-sorted) slice: runs of equal keys are the bincount.
template <typename KeyT>
__global__ void __launch_bounds__(kThreads) mode_count_kernel(const KeyT* __restrict__ keys, int attr_bits,
                                                              const uint32_t* __restrict__ start, int64_t nvox, double threshold,
                                                              long long* __restrict__ mode, long long* __restrict__ mode_count) {
  const int64_t v = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (v >= nvox) return;
  const int64_t a = start[v], b = start[v + 1];
  const double total = (double)(b - a);
  const unsigned long long mask = (1ull << attr_bits) - 1ull;
  long long best_val = 0, best_cnt = 0, above = 0;
  int64_t p = a;
  while (p < b) {
    const KeyT k = keys[p];
    int64_t q = p + 1;
    while (q < b && keys[q] == k) ++q;
    const long long cnt = (long long)(q - p);
    if (cnt > best_cnt) { best_cnt = cnt; best_val = (long long)((unsigned long long)k & mask); }  // first maximum = smallest value (np.argmax)
    if ((double)cnt / total >= threshold) ++above;                              // counts / counts_sum >= percentage
    p = q;
  }
  if (mode) mode[v] = best_val;
  if (mode_count) mode_count[v] = above;
}

// The same from (voxel row << attr_bits | value, weight) runs: sorted distinct keys with the number of points behind each —
// what the owners of a multi-GPU table hold after merging the ranks' runs.  One thread per voxel finds its runs by
// binary search.
__global__ void __launch_bounds__(kThreads) mode_count_runs_kernel(const unsigned long long* __restrict__ run_keys, const long long* __restrict__ run_weight,
                                                                   int64_t nruns, int attr_bits, int64_t nvox, double threshold,
                                                                   long long* __restrict__ mode, long long* __restrict__ mode_count) {
  const int64_t v = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (v >= nvox) return;
  const unsigned long long first = (unsigned long long)v << attr_bits;
  int64_t lo = 0, hi = nruns;
  while (lo < hi) {
    const int64_t mid = (lo + hi) >> 1;
    if (run_keys[mid] < first) lo = mid + 1; else hi = mid;
  }
  const unsigned long long mask = (1ull << attr_bits) - 1ull;
  long long total = 0;
  int64_t e = lo;
  while (e < nruns && (run_keys[e] >> attr_bits) == (unsigned long long)v) total += run_weight[e++];
  long long best_val = 0, best_cnt = 0, above = 0;
  for (int64_t p = lo; p < e; ++p) {
    const long long cnt = run_weight[p];
    if (cnt > best_cnt) { best_cnt = cnt; best_val = (long long)(run_keys[p] & mask); }
    if ((double)cnt / (double)total >= threshold) ++above;
  }
  if (mode) mode[v] = best_val;
  if (mode_count) mode_count[v] = above;
}

// ---- per-voxel statistics of an fp64 attribute (vapc.py:393-406) -----------------------------------------
// One thread per voxel walks the voxel's points in original order (order = idx_sorted) — the order of the reference's slice
// `sorted_data[attr][start:end]` (a stable lexsort by voxel) — and adds them the way numpy's `.sum()` / `.mean()` / `.var()` add a
// 1-D fp64 slice (np_sum.cuh), so that the columns equal the reference's bit for bit:
//   mean = sum / n;  var = pairwise sum of ((x - mean) * (x - mean)) / n  (ddof = 0, numpy/_core/_methods.py:_var);  no FMA anywhere.
__global__ void __launch_bounds__(kThreads) attr_stats_kernel(const double* __restrict__ attr, const uint32_t* __restrict__ order,
                                                              const uint32_t* __restrict__ start, int64_t nvox, double* __restrict__ sum,
                                                              double* __restrict__ mean, double* __restrict__ vmin, double* __restrict__ vmax,
                                                              double* __restrict__ var) {
  const int64_t v = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (v >= nvox) return;
  const int64_t a = start[v], b = start[v + 1];
  const double n = (double)(b - a);
  auto value = [&](int64_t p) { return __ldg(attr + order[p]); };
  const double s = np_pairwise_sum(value, a, b - a);
  const double m = __ddiv_rn(s, n);
  if (sum) sum[v] = s;
  if (mean) mean[v] = m;
  if (vmin || vmax) {
    double lo = INFINITY, hi = -INFINITY;
    bool any_nan = false;
    for (int64_t p = a; p < b; ++p) {
      const double x = value(p);
      any_nan |= (x != x);
      lo = fmin(lo, x);
      hi = fmax(hi, x);
    }
    if (any_nan) lo = hi = s;  // numpy min / max propagate NaN
    if (vmin) vmin[v] = lo;
    if (vmax) vmax[v] = hi;
  }
  if (var) {
    auto sq_dev = [&](int64_t p) {
      const double d = __dsub_rn(value(p), m);
      return __dmul_rn(d, d);
    };
    var[v] = __ddiv_rn(np_pairwise_sum(sq_dev, a, b - a), n);
  }
}

// order-preserving map fp64 -> uint64 (for radix sorting by value); NaN sorts last like numpy
__global__ void __launch_bounds__(kThreads) f64_sort_keys_kernel(const double* __restrict__ attr, int64_t n,
                                                                 unsigned long long* __restrict__ keys) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    const double x = attr[i];
    unsigned long long u = (unsigned long long)__double_as_longlong(x);
    if (x != x) u = ~0ull;
    else u = (u >> 63) ? ~u : (u | 0x8000000000000000ull);
    keys[i] = u;
  }
}

// median per voxel; order = points sorted by (voxel row, value)
__global__ void __launch_bounds__(kThreads) attr_median_kernel(const double* __restrict__ attr, const uint32_t* __restrict__ order,
                                                               const uint32_t* __restrict__ start, int64_t nvox, double* __restrict__ median) {
  const int64_t v = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (v >= nvox) return;
  const int64_t a = start[v], b = start[v + 1], n = b - a;
  const double hi = attr[order[a + n / 2]];
  if (n & 1) {
    median[v] = hi;
  } else {
    const double lo = attr[order[a + n / 2 - 1]];
    median[v] = (lo + hi) / 2.0;  // np.median: mean of the two middle values
  }
  const double last = attr[order[b - 1]];
  if (last != last) median[v] = last;  // numpy: NaN anywhere -> NaN
}

// ---- generic packing of int64 voxel triples --------------------------------------------------------
__global__ void __launch_bounds__(kThreads) pack_triples_kernel(const long long* __restrict__ vx, const long long* __restrict__ vy,
                                                                const long long* __restrict__ vz, int64_t n, Grid g,
                                                                unsigned long long* __restrict__ keys, int* __restrict__ status) {
  int bad = 0;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    const unsigned long long rx = (unsigned long long)(vx[i] - g.minx), ry = (unsigned long long)(vy[i] - g.miny),
                             rz = (unsigned long long)(vz[i] - g.minz);
    const bool out = (g.bx < 64 && (rx >> g.bx)) || (g.by < 64 && (ry >> g.by)) || (g.bz < 64 && (rz >> g.bz));
    bad |= out ? 1 : 0;
    unsigned long long k = rz;
    if (g.bz < 64) k |= ry << g.bz;
    if (g.by + g.bz < 64) k |= rx << (g.by + g.bz);
    keys[i] = out ? kEmpty : k;
  }
  if (bad) atomicOr(status, 1);
}

// ---- open-addressing hash set of packed keys -----------------------------------------------------------
__device__ __forceinline__ unsigned long long hash_slot(unsigned long long k, int log2cap) {
  return (k * 0x9E3779B97F4A7C15ull) >> (64 - log2cap);
}
__global__ void __launch_bounds__(kThreads) mask_build_kernel(const unsigned long long* __restrict__ keys, int64_t n,
                                                              unsigned long long* __restrict__ table, int log2cap) {
  const unsigned long long cap_mask = (1ull << log2cap) - 1ull;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    const unsigned long long k = keys[i];
    if (k == kEmpty) continue;
    unsigned long long s = hash_slot(k, log2cap);
    while (true) {
      const unsigned long long old = atomicCAS(table + s, kEmpty, k);
      if (old == kEmpty || old == k) break;
      s = (s + 1) & cap_mask;
    }
  }
}
__device__ __forceinline__ bool mask_contains(const unsigned long long* __restrict__ table, int log2cap, unsigned long long k) {
  const unsigned long long cap_mask = (1ull << log2cap) - 1ull;
  unsigned long long s = hash_slot(k, log2cap);
  while (true) {
    const unsigned long long t = __ldg(table + s);
    if (t == k) return true;
    if (t == kEmpty) return false;
    s = (s + 1) & cap_mask;
  }
}
// membership of every POINT's voxel: fused voxelise (vapc.py:199-201) + pack in the mask's grid + probe
__global__ void __launch_bounds__(kThreads) mask_lookup_points_kernel(const double* __restrict__ x, const double* __restrict__ y,
                                                                      const double* __restrict__ z, int64_t n, Grid g,
                                                                      const unsigned long long* __restrict__ table, int log2cap,
                                                                      unsigned char* __restrict__ member) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    const unsigned long long rx = (unsigned long long)(voxel_of(x[i], g.ox, g.vs) - g.minx);
    const unsigned long long ry = (unsigned long long)(voxel_of(y[i], g.oy, g.vs) - g.miny);
    const unsigned long long rz = (unsigned long long)(voxel_of(z[i], g.oz, g.vs) - g.minz);
    const bool out = (g.bx < 64 && (rx >> g.bx)) || (g.by < 64 && (ry >> g.by)) || (g.bz < 64 && (rz >> g.bz));
    unsigned long long k = rz;
    if (g.bz < 64) k |= ry << g.bz;
    if (g.by + g.bz < 64) k |= rx << (g.by + g.bz);
    member[i] = (!out && mask_contains(table, log2cap, k)) ? 1 : 0;
  }
}
__global__ void __launch_bounds__(kThreads) mask_lookup_keys_kernel(const unsigned long long* __restrict__ keys, int64_t n,
                                                                    const unsigned long long* __restrict__ table, int log2cap,
                                                                    unsigned char* __restrict__ member) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    const unsigned long long k = keys[i];
    member[i] = (k != kEmpty && mask_contains(table, log2cap, k)) ? 1 : 0;
  }
}

// ---- voxel buffer: every triple + all (2b+1)^3 offsets, voxel-major, offsets in 'ij' meshgrid order ---
__global__ void __launch_bounds__(kThreads) buffer_expand_kernel(const long long* __restrict__ vx, const long long* __restrict__ vy,
                                                                 const long long* __restrict__ vz, int64_t n, int b,
                                                                 long long* __restrict__ out) {
  const int w = 2 * b + 1;
  const int64_t per = (int64_t)w * w * w, total = n * per;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
    const int64_t v = i / per;
    const int o = (int)(i - v * per);
    const int dz = o % w - b, dy = (o / w) % w - b, dx = o / (w * w) - b;
    out[i] = vx[v] + dx;
    out[total + i] = vy[v] + dy;
    out[2 * total + i] = vz[v] + dz;
  }
}

inline int ilog2_ceil(int64_t v) {
  int b = 0;
  while (((int64_t)1 << b) < v) ++b;
  return b;
}
}  // namespace

extern "C" int vapc_mode_keys(const uint32_t* point2voxel, const double* attr, int64_t n, int32_t attr_bits, int32_t key_bytes, void* keys,
                              int32_t* status, void* stream) {
  if (n < 0 || attr_bits < 1 || attr_bits > 32 || (key_bytes != 4 && key_bytes != 8))
    return vapc_set_error(VAPC_E_INVALID, "vapc_mode_keys: bad arguments (attr_bits in 1..32, key_bytes 4 or 8)");
  if (n == 0) return VAPC_OK;
  if (key_bytes == 4)
    VAPC_LAUNCH(mode_keys_kernel<uint32_t>, grid_for(n), kThreads, 0, as_stream(stream), point2voxel, attr, n, attr_bits, static_cast<uint32_t*>(keys), status);
  else
    VAPC_LAUNCH(mode_keys_kernel<unsigned long long>, grid_for(n), kThreads, 0, as_stream(stream), point2voxel, attr, n, attr_bits,
                static_cast<unsigned long long*>(keys), status);
  VAPC_CHECK_LAUNCH();
  return VAPC_OK;
}

extern "C" int vapc_mode_count(const void* keys_sorted, int64_t n, int32_t attr_bits, int32_t key_bytes, const uint32_t* start, int64_t nvoxels,
                               double threshold, int64_t* mode, int64_t* mode_count, void* stream) {
  if (n < 0 || nvoxels < 0 || (key_bytes != 4 && key_bytes != 8)) return vapc_set_error(VAPC_E_INVALID, "vapc_mode_count: bad arguments");
  if (nvoxels == 0) return VAPC_OK;
  const unsigned blocks = (unsigned)((nvoxels + kThreads - 1) / kThreads);
  if (key_bytes == 4)
    VAPC_LAUNCH(mode_count_kernel<uint32_t>, blocks, kThreads, 0, as_stream(stream), static_cast<const uint32_t*>(keys_sorted), attr_bits, start, nvoxels,
                threshold, reinterpret_cast<long long*>(mode), reinterpret_cast<long long*>(mode_count));
  else
    VAPC_LAUNCH(mode_count_kernel<unsigned long long>, blocks, kThreads, 0, as_stream(stream), static_cast<const unsigned long long*>(keys_sorted), attr_bits,
                start, nvoxels, threshold, reinterpret_cast<long long*>(mode), reinterpret_cast<long long*>(mode_count));
  VAPC_CHECK_LAUNCH();
  return VAPC_OK;
}

extern "C" int vapc_mode_count_runs(const uint64_t* run_keys, const int64_t* run_weight, int64_t nruns, int32_t attr_bits, int64_t nvoxels,
                                    double threshold, int64_t* mode, int64_t* mode_count, void* stream) {
  if (nruns < 0 || nvoxels < 0 || attr_bits < 1 || attr_bits > 40) return vapc_set_error(VAPC_E_INVALID, "vapc_mode_count_runs: bad arguments");
  if (nvoxels == 0) return VAPC_OK;
  if (nruns > 0 && (!run_keys || !run_weight)) return vapc_set_error(VAPC_E_INVALID, "vapc_mode_count_runs: null buffer");
  VAPC_LAUNCH(mode_count_runs_kernel, (unsigned)((nvoxels + kThreads - 1) / kThreads), kThreads, 0, as_stream(stream),
              reinterpret_cast<const unsigned long long*>(run_keys), reinterpret_cast<const long long*>(run_weight), nruns, attr_bits, nvoxels, threshold,
              reinterpret_cast<long long*>(mode), reinterpret_cast<long long*>(mode_count));
  VAPC_CHECK_LAUNCH();
  return VAPC_OK;
}

extern "C" int vapc_attr_stats(const double* attr, const uint32_t* order, const uint32_t* start, int64_t nvoxels, double* sum,
                               double* mean, double* vmin, double* vmax, double* var, void* stream) {
  if (nvoxels < 0) return vapc_set_error(VAPC_E_INVALID, "vapc_attr_stats: bad arguments");
  if (nvoxels == 0) return VAPC_OK;
  VAPC_LAUNCH(attr_stats_kernel, (unsigned)((nvoxels + kThreads - 1) / kThreads), kThreads, 0, as_stream(stream), attr, order, start, nvoxels, sum, mean,
                                                                                                       vmin, vmax, var);
  VAPC_CHECK_LAUNCH();
  return VAPC_OK;
}
extern "C" int vapc_f64_sort_keys(const double* attr, int64_t n, uint64_t* keys, void* stream) {
  if (n < 0) return vapc_set_error(VAPC_E_INVALID, "vapc_f64_sort_keys: n < 0");
  if (n == 0) return VAPC_OK;
  VAPC_LAUNCH(f64_sort_keys_kernel, grid_for(n), kThreads, 0, as_stream(stream), attr, n, reinterpret_cast<unsigned long long*>(keys));
  VAPC_CHECK_LAUNCH();
  return VAPC_OK;
}
extern "C" int vapc_attr_median(const double* attr, const uint32_t* order, const uint32_t* start, int64_t nvoxels, double* median,
                                void* stream) {
  if (nvoxels < 0) return vapc_set_error(VAPC_E_INVALID, "vapc_attr_median: bad arguments");
  if (nvoxels == 0) return VAPC_OK;
  VAPC_LAUNCH(attr_median_kernel, (unsigned)((nvoxels + kThreads - 1) / kThreads), kThreads, 0, as_stream(stream), attr, order, start, nvoxels, median);
  VAPC_CHECK_LAUNCH();
  return VAPC_OK;
}

extern "C" int vapc_pack_triples_i64(const int64_t* vx, const int64_t* vy, const int64_t* vz, int64_t n, const vapc_grid_t* grid_host,
                                     uint64_t* keys, int32_t* status, void* stream) {
  if (!grid_host || n < 0) return vapc_set_error(VAPC_E_INVALID, "vapc_pack_triples_i64: bad arguments");
  if (grid_host->bits[0] + grid_host->bits[1] + grid_host->bits[2] > 63)
    return vapc_set_error(VAPC_E_KEYSPACE, "vapc_pack_triples_i64: needs at most 63 key bits");
  if (n == 0) return VAPC_OK;
  VAPC_LAUNCH(pack_triples_kernel, grid_for(n), kThreads, 0, as_stream(stream), reinterpret_cast<const long long*>(vx),
      reinterpret_cast<const long long*>(vy), reinterpret_cast<const long long*>(vz), n, to_device_grid(grid_host),
      reinterpret_cast<unsigned long long*>(keys), status);
  VAPC_CHECK_LAUNCH();
  return VAPC_OK;
}

extern "C" size_t vapc_mask_table_bytes(int64_t nmask, int64_t* capacity_host) {
  const int lg = ilog2_ceil((nmask < 1 ? 1 : nmask) * 2);
  const int64_t cap = (int64_t)1 << (lg < 4 ? 4 : lg);
  if (capacity_host) *capacity_host = cap;
  return (size_t)cap * sizeof(uint64_t);
}

extern "C" int vapc_mask_build(const uint64_t* mask_keys, int64_t nmask, void* table, int64_t capacity, void* stream) {
  if (nmask < 0 || capacity < 16 || (capacity & (capacity - 1)) || capacity < 2 * nmask || !table)
    return vapc_set_error(VAPC_E_INVALID, "vapc_mask_build: capacity must be a power of two >= max(16, 2 * nmask)");
  VAPC_RETURN_IF_CUDA(cudaMemsetAsync(table, 0xff, (size_t)capacity * sizeof(uint64_t), as_stream(stream)));
  if (nmask == 0) return VAPC_OK;
  VAPC_LAUNCH(mask_build_kernel, grid_for(nmask, 1), kThreads, 0, as_stream(stream), reinterpret_cast<const unsigned long long*>(mask_keys), nmask,
                                                                            static_cast<unsigned long long*>(table), ilog2_ceil(capacity));
  VAPC_CHECK_LAUNCH();
  return VAPC_OK;
}

extern "C" int vapc_mask_lookup_points(const double* x, const double* y, const double* z, int64_t n, const vapc_grid_t* grid_host,
                                       const void* table, int64_t capacity, uint8_t* member, void* stream) {
  if (!grid_host || n < 0 || capacity < 16 || (capacity & (capacity - 1))) return vapc_set_error(VAPC_E_INVALID, "vapc_mask_lookup_points: bad arguments");
  if (n == 0) return VAPC_OK;
  VAPC_LAUNCH(mask_lookup_points_kernel, grid_for(n), kThreads, 0, as_stream(stream), x, y, z, n, to_device_grid(grid_host),
      static_cast<const unsigned long long*>(table), ilog2_ceil(capacity), member);
  VAPC_CHECK_LAUNCH();
  return VAPC_OK;
}

extern "C" int vapc_mask_lookup_keys(const uint64_t* keys, int64_t n, const void* table, int64_t capacity, uint8_t* member, void* stream) {
  if (n < 0 || capacity < 16 || (capacity & (capacity - 1))) return vapc_set_error(VAPC_E_INVALID, "vapc_mask_lookup_keys: bad arguments");
  if (n == 0) return VAPC_OK;
  VAPC_LAUNCH(mask_lookup_keys_kernel, grid_for(n), kThreads, 0, as_stream(stream), reinterpret_cast<const unsigned long long*>(keys), n,
      static_cast<const unsigned long long*>(table), ilog2_ceil(capacity), member);
  VAPC_CHECK_LAUNCH();
  return VAPC_OK;
}

extern "C" int vapc_buffer_expand(const int64_t* vx, const int64_t* vy, const int64_t* vz, int64_t n, int32_t buffer_size, int64_t* out3,
                                  void* stream) {
  if (n < 0 || buffer_size < 0 || buffer_size > 64) return vapc_set_error(VAPC_E_INVALID, "vapc_buffer_expand: bad arguments");
  if (n == 0) return VAPC_OK;
  const int w = 2 * buffer_size + 1;
  VAPC_LAUNCH(buffer_expand_kernel, grid_for(n * w * w * w), kThreads, 0, as_stream(stream), reinterpret_cast<const long long*>(vx),
      reinterpret_cast<const long long*>(vy), reinterpret_cast<const long long*>(vz), n, buffer_size, reinterpret_cast<long long*>(out3));
  VAPC_CHECK_LAUNCH();
  return VAPC_OK;
}

// histogram of the top hist_bits of the keys (multi-GPU splitter selection)
namespace {
template <typename KeyT>
__global__ void __launch_bounds__(kThreads) key_hist_kernel(const KeyT* __restrict__ keys, int64_t n, int shift, uint32_t* __restrict__ hist) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
    atomicAdd(hist + (unsigned)((unsigned long long)keys[i] >> shift), 1u);
}
}  // namespace
extern "C" int vapc_key_histogram(const void* keys, int64_t n, int32_t key_bytes, int32_t total_bits, int32_t hist_bits, uint32_t* hist,
                                  void* stream) {
  if (n < 0 || hist_bits < 1 || hist_bits > 24 || total_bits < 0 || total_bits > 64)
    return vapc_set_error(VAPC_E_INVALID, "vapc_key_histogram: bad arguments");
  VAPC_RETURN_IF_CUDA(cudaMemsetAsync(hist, 0, sizeof(uint32_t) << hist_bits, as_stream(stream)));
  if (n == 0) return VAPC_OK;
  const int shift = total_bits > hist_bits ? total_bits - hist_bits : 0;
  if (key_bytes == 4)
    VAPC_LAUNCH(key_hist_kernel<uint32_t>, grid_for(n), kThreads, 0, as_stream(stream), static_cast<const uint32_t*>(keys), n, shift, hist);
  else if (key_bytes == 8)
    VAPC_LAUNCH(key_hist_kernel<unsigned long long>, grid_for(n), kThreads, 0, as_stream(stream), static_cast<const unsigned long long*>(keys), n, shift, hist);
  else
    return vapc_set_error(VAPC_E_INVALID, "vapc_key_histogram: key_bytes must be 4 or 8");
  VAPC_CHECK_LAUNCH();
  return VAPC_OK;
}

// ---- multi-GPU exchange planning on the device --------------------------------------------------------------------------
// hist_all[r][b]: rank r's histogram of the leading key bits of its partial voxel rows (all-gathered).  One CTA finds the
// world - 1 splitters that balance the GLOBAL number of rows per key range — b_g = (first bin whose cumulative count reaches
// total * g / world) + 1, exactly vapc_b200.dist.plan_splitters — and, because splitters lie on bin boundaries, every rank's
// send counts to every destination as sums over bin ranges.  out = [matrix world x world (row r = rank r's send counts)]
// [thresholds world - 1] as int64, so that the host learns everything with one copy.
namespace {
constexpr int kPlanThreads = 1024;
constexpr int kMaxWorld = 64;
// block-wide inclusive scan of one u64 per thread; returns the inclusive prefix, *total = block sum.  Ends with a barrier.
__device__ __forceinline__ unsigned long long plan_block_scan(unsigned long long v, unsigned long long* warp_sums, unsigned long long* total) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  unsigned long long inc = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const unsigned long long t = __shfl_up_sync(0xffffffffu, inc, o);
    if (lane >= o) inc += t;
  }
  if (lane == 31) warp_sums[warp] = inc;
  __syncthreads();
  if (warp == 0) {
    const unsigned long long w = warp_sums[lane];
    unsigned long long winc = w;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const unsigned long long t = __shfl_up_sync(0xffffffffu, winc, o);
      if (lane >= o) winc += t;
    }
    warp_sums[lane] = winc - w;
    if (lane == 31) warp_sums[32] = winc;
  }
  __syncthreads();
  const unsigned long long res = inc + warp_sums[warp];
  *total = warp_sums[32];
  __syncthreads();
  return res;
}

__global__ void __launch_bounds__(kPlanThreads) plan_exchange_kernel(const uint32_t* __restrict__ hist_all, int world, int nbins, int shift,
                                                                     long long* __restrict__ out) {
  __shared__ unsigned long long warp_sums[33];
  __shared__ unsigned long long matrix[kMaxWorld * kMaxWorld];
  __shared__ int edges[kMaxWorld + 1];
  const int tid = threadIdx.x;
  for (int i = tid; i < world * world; i += kPlanThreads) matrix[i] = 0;
  if (tid == 0) { edges[0] = 0; edges[world] = nbins; }
  // pass 1: total number of rows (bins are visited 1024 at a time, thread t takes bin round * 1024 + t: coalesced)
  unsigned long long mine = 0;
  for (int b = tid; b < nbins; b += kPlanThreads)
    for (int r = 0; r < world; ++r) mine += hist_all[(int64_t)r * nbins + b];
  unsigned long long total;
  plan_block_scan(mine, warp_sums, &total);
  // pass 2: running cumulative count; the thread whose bin is the first to reach a target sets that splitter
  unsigned long long carry = 0;
  for (int b0 = 0; b0 < nbins; b0 += kPlanThreads) {
    const int b = b0 + tid;
    unsigned long long gb = 0;
    if (b < nbins)
      for (int r = 0; r < world; ++r) gb += hist_all[(int64_t)r * nbins + b];
    unsigned long long round_total;
    const unsigned long long cum = carry + plan_block_scan(gb, warp_sums, &round_total);
    if (b < nbins) {
      const unsigned long long before = cum - gb;
      for (int g = 1; g < world; ++g) {
        const double target = (double)total * (double)g / (double)world;
        if ((double)cum >= target && (b == 0 || !((double)before >= target))) edges[g] = min(b + 1, nbins);
      }
    }
    carry += round_total;
  }
  __syncthreads();
  if (tid == 0)  // monotone (cummax), as the host planner does
    for (int g = 1; g < world; ++g) edges[g] = max(edges[g], edges[g - 1]);
  __syncthreads();
  // pass 3: send counts = every rank's rows in the bins of every destination.  A warp walks a contiguous range of bins 32 at
  // a time (coalesced); the destination changes at most world - 1 times over all bins, so lanes accumulate in a register and
  // only flush (warp reduction + one shared-memory atomic) when it does.
  {
    const int lane = tid & 31, warp = tid >> 5, nwarps = kPlanThreads / 32;
    const int per_warp = ((nbins + nwarps - 1) / nwarps + 31) / 32 * 32;
    const int w0 = warp * per_warp, w1 = min(nbins, w0 + per_warp);
    for (int r = 0; r < world; ++r) {
      const uint32_t* h = hist_all + (int64_t)r * nbins;
      int g = 0;
      unsigned long long acc = 0;
      for (int s0 = w0; s0 < w1; s0 += 32) {
        const int b = s0 + lane;
        const unsigned long long c = b < w1 ? h[b] : 0u;
        const int last = min(s0 + 31, w1 - 1);
        while (s0 >= edges[g + 1]) {  // the whole step lies beyond destination g: flush and advance (warp-uniform)
          unsigned long long t = acc;
#pragma unroll
          for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
          if (lane == 0 && t) atomicAdd(&matrix[r * world + g], t);
          acc = 0;
          ++g;
        }
        if (last < edges[g + 1]) {
          acc += c;  // every bin of the step belongs to g
        } else if (b < w1 && c) {  // a step that straddles a splitter: per-lane destination, direct atomics
          int gg = g;
          while (b >= edges[gg + 1]) ++gg;
          atomicAdd(&matrix[r * world + gg], c);
        }
      }
      unsigned long long t = acc;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
      if (lane == 0 && t) atomicAdd(&matrix[r * world + g], t);
    }
  }
  __syncthreads();
  for (int i = tid; i < world * world; i += kPlanThreads) out[i] = (long long)matrix[i];
  for (int g = tid + 1; g < world; g += kPlanThreads) out[world * world + g - 1] = (long long)edges[g] << shift;
}

// keys of the rows a rank merges: its own rows [own_first, own_first + n_own) of the global-key array, then the keys of the
// received exchange rows (first word of each 11-word row), as 32- or 64-bit sort keys
template <typename KeyT>
__global__ void __launch_bounds__(kThreads) merge_keys_kernel(const unsigned long long* __restrict__ gkeys, int64_t own_first, int64_t n_own,
                                                              const double* __restrict__ rows, int64_t n_in, KeyT* __restrict__ out) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x, n = n_own + n_in;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
    out[i] = (KeyT)(i < n_own ? gkeys[own_first + i] : (unsigned long long)__double_as_longlong(rows[(i - n_own) * 11]));
}
}  // namespace

extern "C" int vapc_plan_exchange(const uint32_t* hist_all, int32_t world, int32_t hist_bits, int32_t global_bits, int64_t* out, void* stream) {
  if (!hist_all || !out || world < 1 || world > kMaxWorld || hist_bits < 1 || hist_bits > 24)
    return vapc_set_error(VAPC_E_INVALID, "vapc_plan_exchange: bad arguments (1 <= world <= %d)", kMaxWorld);
  const int shift = global_bits > hist_bits ? global_bits - hist_bits : 0;
  VAPC_LAUNCH(plan_exchange_kernel, 1, kPlanThreads, 0, as_stream(stream), hist_all, world, 1 << hist_bits, shift, reinterpret_cast<long long*>(out));
  VAPC_CHECK_LAUNCH();
  return VAPC_OK;
}

extern "C" int vapc_merge_keys(const uint64_t* global_keys, int64_t own_first, int64_t n_own, const double* rows, int64_t n_in, int32_t key_bytes,
                               void* keys_out, void* stream) {
  if (n_own < 0 || n_in < 0 || own_first < 0) return vapc_set_error(VAPC_E_INVALID, "vapc_merge_keys: bad arguments");
  if (n_own + n_in == 0) return VAPC_OK;
  if (!keys_out || (n_own && !global_keys) || (n_in && !rows)) return vapc_set_error(VAPC_E_INVALID, "vapc_merge_keys: null buffer");
  const unsigned long long* gk = reinterpret_cast<const unsigned long long*>(global_keys);
  if (key_bytes == 4)
    VAPC_LAUNCH(merge_keys_kernel<uint32_t>, grid_for(n_own + n_in), kThreads, 0, as_stream(stream), gk, own_first, n_own, rows, n_in, static_cast<uint32_t*>(keys_out));
  else if (key_bytes == 8)
    VAPC_LAUNCH(merge_keys_kernel<unsigned long long>, grid_for(n_own + n_in), kThreads, 0, as_stream(stream), gk, own_first, n_own, rows, n_in,
                static_cast<unsigned long long*>(keys_out));
  else
    return vapc_set_error(VAPC_E_INVALID, "vapc_merge_keys: key_bytes must be 4 or 8");
  VAPC_CHECK_LAUNCH();
  return VAPC_OK;
}

// host mirror of the planning kernel's arithmetic for the CPU tests (no GPU needed)
extern "C" int vapc_selftest_plan_exchange_host(const uint32_t* hist_all_host, int32_t world, int32_t nbins, int32_t shift, int64_t* out_host) {
  std::vector<unsigned long long> cum(nbins);
  unsigned long long run = 0;
  for (int b = 0; b < nbins; ++b) {
    for (int r = 0; r < world; ++r) run += hist_all_host[(int64_t)r * nbins + b];
    cum[b] = run;
  }
  std::vector<int> edges(world + 1, 0);
  edges[world] = nbins;
  for (int g = 1; g < world; ++g) {
    const double target = (double)run * (double)g / (double)world;
    int b = 0;
    while (b < nbins && !((double)cum[b] >= target)) ++b;
    edges[g] = std::max(std::min(b + 1, nbins), edges[g - 1]);
  }
  for (int r = 0; r < world; ++r)
    for (int g = 0; g < world; ++g) {
      unsigned long long acc = 0;
      for (int b = edges[g]; b < edges[g + 1]; ++b) acc += hist_all_host[(int64_t)r * nbins + b];
      out_host[r * world + g] = (long long)acc;
    }
  for (int g = 1; g < world; ++g) out_host[world * world + g - 1] = (long long)edges[g] << shift;
  return VAPC_OK;
}
