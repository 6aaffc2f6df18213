// Per-voxel outputs from the (count, mean, M2) table: density, centroid, std, covariance,
// eigenvalues (cyclic Jacobi on the symmetric 3x3), the 8 eigenvalue features and the 3
// normalised eigenvalues.  One thread per voxel; reads 4 + 72 (+ key) bytes, writes up to 192.
//
// Reference formulas: vapc.py:740 (density), 843-844 (centroid), 877-878 (std, ddof = 1),
// 977-1006 (covariance, NaN when n <= 1), 1055-1065 (eigenvalues of the ddof=1 covariance, sorted
// descending, NaN when n == 1), 1112-1145 (features).
#include <math.h>

#include "common.cuh"
#include "math3.cuh"
#include "stats_row.cuh"

namespace {
constexpr int kThreads = 128;
template <typename KeyT>
__global__ void __launch_bounds__(kThreads) voxel_stats_kernel(const KeyT* __restrict__ keys, const uint32_t* __restrict__ count,
                                                               const double* __restrict__ mean3, const double* __restrict__ m2,
                                                               int64_t nvox, Grid g, StatsOut o) {
  const int64_t v = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (v >= nvox) return;
  double mean[3] = {0, 0, 0}, M[6] = {0, 0, 0, 0, 0, 0};
  if (o.cog3) {
#pragma unroll
    for (int k = 0; k < 3; ++k) mean[k] = mean3[(int64_t)k * nvox + v];
  }
  if (o.needs_m2()) {
#pragma unroll
    for (int k = 0; k < 6; ++k) M[k] = m2[(int64_t)k * nvox + v];
  }
  emit_voxel_stats((unsigned long long)keys[v], count[v], mean, M, g, v, o);
}

// out[i] = table[index[i]]
template <typename T>
__global__ void __launch_bounds__(256) gather_kernel(const T* __restrict__ table, const uint32_t* __restrict__ index, int64_t n,
                                                     T* __restrict__ out) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) out[i] = __ldg(table + __ldcs(index + i));
}

__global__ void __launch_bounds__(256) point_distance_kernel(const double* __restrict__ x, const double* __restrict__ y,
                                                             const double* __restrict__ z, int64_t n,
                                                             const uint32_t* __restrict__ p2v, const double* __restrict__ cog3,
                                                             int64_t nvox, double* __restrict__ dist) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    const uint32_t v = p2v[i];
    const double dx = __dsub_rn(x[i], cog3[v]), dy = __dsub_rn(y[i], cog3[nvox + v]), dz = __dsub_rn(z[i], cog3[2 * nvox + v]);
    dist[i] = __dsqrt_rn(__dadd_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)), __dmul_rn(dz, dz)));
  }
}

inline int grid_for(int64_t n, int threads, int per_thread = 1) {
  int64_t b = (n + (int64_t)threads * per_thread - 1) / ((int64_t)threads * per_thread);
  const int64_t cap = (int64_t)kNumSMs * 32;
  if (b > cap) b = cap;
  return (int)(b < 1 ? 1 : b);
}
}  // namespace

extern "C" int vapc_voxel_stats_axes(const void* voxel_keys, const uint32_t* count, const double* mean3, const double* m2,
                                     int64_t nvoxels, const vapc_grid_t* grid_host, const int32_t* axis_order_host, double* density,
                                     double* cog3, double* std3, double* cov6, double* eig3, double* eign3, double* feat8, void* stream) {
  if (!grid_host || nvoxels < 0) return vapc_set_error(VAPC_E_INVALID, "vapc_voxel_stats: bad arguments");
  StatsOut o{density, cog3, std3, cov6, eig3, eign3, feat8, nvoxels, {}};
  if (!stats_axes_from_order(axis_order_host, &o.ax)) return vapc_set_error(VAPC_E_INVALID, "vapc_voxel_stats_axes: axis_order must be a permutation of 0, 1, 2");
  if (nvoxels == 0) return VAPC_OK;
  if (!voxel_keys || !count || !mean3 || !m2) return vapc_set_error(VAPC_E_INVALID, "vapc_voxel_stats: null input");
  const Grid g = to_device_grid(grid_host);
  const unsigned blocks = (unsigned)((nvoxels + kThreads - 1) / kThreads);
  if (grid_host->key_bytes == 4)
    VAPC_LAUNCH(voxel_stats_kernel<uint32_t>, blocks, kThreads, 0, as_stream(stream), static_cast<const uint32_t*>(voxel_keys), count, mean3, m2,
                                                                             nvoxels, g, o);
  else
    VAPC_LAUNCH(voxel_stats_kernel<unsigned long long>, blocks, kThreads, 0, as_stream(stream), 
        static_cast<const unsigned long long*>(voxel_keys), count, mean3, m2, nvoxels, g, o);
  VAPC_CHECK_LAUNCH();
  return VAPC_OK;
}

extern "C" int vapc_voxel_stats(const void* voxel_keys, const uint32_t* count, const double* mean3, const double* m2,
                                int64_t nvoxels, const vapc_grid_t* grid_host, double* density, double* cog3, double* std3,
                                double* cov6, double* eig3, double* eign3, double* feat8, void* stream) {
  return vapc_voxel_stats_axes(voxel_keys, count, mean3, m2, nvoxels, grid_host, nullptr, density, cog3, std3, cov6, eig3, eign3, feat8, stream);
}

namespace {
// The reference's OWN covariance columns, rounding for rounding (vapc.py:947-1006): products X*X, X*Y, ... rounded to fp64 per point,
// per-voxel sums the way pandas' groupby sum forms them — Kahan-compensated, rows in frame order (inside a voxel the sorted
// positions ARE in original row order: the sort is stable) — then (S_ab - S_a S_b / n) / (n - 1) with one rounding per operation.
// One thread per voxel: the recurrence is sequential by definition.  No FMA anywhere.
struct Kahan {
  double s = 0.0, c = 0.0;
  __device__ __forceinline__ void add(double v) {
    const double y = __dsub_rn(v, c);
    const double t = __dadd_rn(s, y);
    c = __dsub_rn(__dsub_rn(t, s), y);
    if (c != c) c = 0.0;  // pandas resets a NaN compensation (infinite sums)
    s = t;
  }
};
__global__ void __launch_bounds__(kThreads) cov_reference_kernel(const double4* __restrict__ xyzw, const uint32_t* __restrict__ pos_sorted,
                                                                 const uint32_t* __restrict__ start, int64_t nvox, double* __restrict__ cov6) {
  const int64_t v = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (v >= nvox) return;
  const uint32_t a = start[v], b = start[v + 1];
  Kahan sx, sy, sz, sxx, syy, szz, sxy, sxz, syz;
  for (uint32_t p = a; p < b; ++p) {
    const double4 r = ld_record(xyzw + pos_sorted[p]);
    sx.add(r.x); sy.add(r.y); sz.add(r.z);
    sxx.add(__dmul_rn(r.x, r.x)); syy.add(__dmul_rn(r.y, r.y)); szz.add(__dmul_rn(r.z, r.z));
    sxy.add(__dmul_rn(r.x, r.y)); sxz.add(__dmul_rn(r.x, r.z)); syz.add(__dmul_rn(r.y, r.z));
  }
  const double n = (double)(b - a), n1 = (double)((long long)(b - a) - 1);
  auto cov = [&](double sab, double sa, double sb) {
    return b - a > 1 ? __ddiv_rn(__dsub_rn(sab, __ddiv_rn(__dmul_rn(sa, sb), n)), n1) : __longlong_as_double(0x7ff8000000000000LL);
  };
  cov6[0 * nvox + v] = cov(sxx.s, sx.s, sx.s);
  cov6[1 * nvox + v] = cov(sxy.s, sx.s, sy.s);
  cov6[2 * nvox + v] = cov(sxz.s, sx.s, sz.s);
  cov6[3 * nvox + v] = cov(syy.s, sy.s, sy.s);
  cov6[4 * nvox + v] = cov(syz.s, sy.s, sz.s);
  cov6[5 * nvox + v] = cov(szz.s, sz.s, sz.s);
}
// The reference's OWN centroid columns, rounding for rounding (vapc.py:843-850): pandas' groupby mean = the Kahan-compensated sum of
// the voxel's X (Y, Z) in frame order, divided by the count.  Equal to the exact mean to the last bit or two — but the reference selects
// its closest-to-centroid rows by `distance == min_distance` on distances to THIS centroid, so ties and near-ties (duplicate points,
// points mirrored about the centroid) only fall the same way when the centroid is the same double.
__global__ void __launch_bounds__(kThreads) cog_reference_kernel(const double4* __restrict__ xyzw, const uint32_t* __restrict__ pos_sorted,
                                                                 const uint32_t* __restrict__ start, int64_t nvox, double* __restrict__ cog3) {
  const int64_t v = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (v >= nvox) return;
  const uint32_t a = start[v], b = start[v + 1];
  Kahan sx, sy, sz;
  for (uint32_t p = a; p < b; ++p) {
    const double4 r = ld_record(xyzw + pos_sorted[p]);
    sx.add(r.x); sy.add(r.y); sz.add(r.z);
  }
  const double n = (double)(b - a);
  cog3[0 * nvox + v] = __ddiv_rn(sx.s, n);
  cog3[1 * nvox + v] = __ddiv_rn(sy.s, n);
  cog3[2 * nvox + v] = __ddiv_rn(sz.s, n);
}
// The reference's OWN std_* columns, rounding for rounding (vapc.py:873-880): pandas' groupby std = Welford's recurrence over the voxel's
// values in frame order (mean += (v - mean) / k; M2 += (v - mean_new) * (v - mean_old)), M2 / (n - 1), square root.  At UTM-sized
// coordinates the recurrence loses ~1e-7 of the value (the reason the default std comes from the shifted moments); this entry point is
// for callers who diff against the reference's column.  NaN for single-point voxels.
__global__ void __launch_bounds__(kThreads) std_reference_kernel(const double4* __restrict__ xyzw, const uint32_t* __restrict__ pos_sorted,
                                                                 const uint32_t* __restrict__ start, int64_t nvox, double* __restrict__ std3) {
  const int64_t v = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (v >= nvox) return;
  const uint32_t a = start[v], b = start[v + 1];
  double mean[3] = {0.0, 0.0, 0.0}, m2[3] = {0.0, 0.0, 0.0};
  double k = 0.0;
  for (uint32_t p = a; p < b; ++p) {
    const double4 r = ld_record(xyzw + pos_sorted[p]);
    const double val[3] = {r.x, r.y, r.z};
    k = __dadd_rn(k, 1.0);
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      const double old = mean[c];
      mean[c] = __dadd_rn(old, __ddiv_rn(__dsub_rn(val[c], old), k));
      m2[c] = __dadd_rn(m2[c], __dmul_rn(__dsub_rn(val[c], mean[c]), __dsub_rn(val[c], old)));
    }
  }
  const double n1 = (double)((long long)(b - a) - 1);
#pragma unroll
  for (int c = 0; c < 3; ++c)
    std3[c * nvox + v] = b - a > 1 ? __dsqrt_rn(__ddiv_rn(m2[c], n1)) : __longlong_as_double(0x7ff8000000000000LL);
}
}  // namespace

extern "C" int vapc_std_reference_formula(const double* xyzw, const uint32_t* pos_sorted, const uint32_t* start, int64_t nvoxels, double* std3,
                                          void* stream) {
  if (nvoxels < 0) return vapc_set_error(VAPC_E_INVALID, "vapc_std_reference_formula: bad arguments");
  if (nvoxels == 0) return VAPC_OK;
  if (!xyzw || !pos_sorted || !start || !std3) return vapc_set_error(VAPC_E_INVALID, "vapc_std_reference_formula: null buffer");
  if (reinterpret_cast<uintptr_t>(xyzw) & 31u) return vapc_set_error(VAPC_E_INVALID, "vapc_std_reference_formula: xyzw must be 32-byte aligned");
  VAPC_LAUNCH(std_reference_kernel, (unsigned)((nvoxels + kThreads - 1) / kThreads), kThreads, 0, as_stream(stream),
              reinterpret_cast<const double4*>(xyzw), pos_sorted, start, nvoxels, std3);
  VAPC_CHECK_LAUNCH();
  return VAPC_OK;
}

extern "C" int vapc_cog_reference_formula(const double* xyzw, const uint32_t* pos_sorted, const uint32_t* start, int64_t nvoxels, double* cog3,
                                          void* stream) {
  if (nvoxels < 0) return vapc_set_error(VAPC_E_INVALID, "vapc_cog_reference_formula: bad arguments");
  if (nvoxels == 0) return VAPC_OK;
  if (!xyzw || !pos_sorted || !start || !cog3) return vapc_set_error(VAPC_E_INVALID, "vapc_cog_reference_formula: null buffer");
  if (reinterpret_cast<uintptr_t>(xyzw) & 31u) return vapc_set_error(VAPC_E_INVALID, "vapc_cog_reference_formula: xyzw must be 32-byte aligned");
  VAPC_LAUNCH(cog_reference_kernel, (unsigned)((nvoxels + kThreads - 1) / kThreads), kThreads, 0, as_stream(stream),
              reinterpret_cast<const double4*>(xyzw), pos_sorted, start, nvoxels, cog3);
  VAPC_CHECK_LAUNCH();
  return VAPC_OK;
}

extern "C" int vapc_cov_reference_formula(const double* xyzw, const uint32_t* pos_sorted, const uint32_t* start, int64_t nvoxels, double* cov6,
                                          void* stream) {
  if (nvoxels < 0) return vapc_set_error(VAPC_E_INVALID, "vapc_cov_reference_formula: bad arguments");
  if (nvoxels == 0) return VAPC_OK;
  if (!xyzw || !pos_sorted || !start || !cov6) return vapc_set_error(VAPC_E_INVALID, "vapc_cov_reference_formula: null buffer");
  if (reinterpret_cast<uintptr_t>(xyzw) & 31u) return vapc_set_error(VAPC_E_INVALID, "vapc_cov_reference_formula: xyzw must be 32-byte aligned");
  VAPC_LAUNCH(cov_reference_kernel, (unsigned)((nvoxels + kThreads - 1) / kThreads), kThreads, 0, as_stream(stream),
              reinterpret_cast<const double4*>(xyzw), pos_sorted, start, nvoxels, cov6);
  VAPC_CHECK_LAUNCH();
  return VAPC_OK;
}

extern "C" int vapc_gather_f64(const double* table, const uint32_t* index, int64_t n, double* out, void* stream) {
  if (n < 0) return vapc_set_error(VAPC_E_INVALID, "vapc_gather_f64: n < 0");
  if (n == 0) return VAPC_OK;
  VAPC_LAUNCH(gather_kernel<double>, grid_for(n, 256, 4), 256, 0, as_stream(stream), table, index, n, out);
  VAPC_CHECK_LAUNCH();
  return VAPC_OK;
}
extern "C" int vapc_gather_u32(const uint32_t* table, const uint32_t* index, int64_t n, uint32_t* out, void* stream) {
  if (n < 0) return vapc_set_error(VAPC_E_INVALID, "vapc_gather_u32: n < 0");
  if (n == 0) return VAPC_OK;
  VAPC_LAUNCH(gather_kernel<uint32_t>, grid_for(n, 256, 4), 256, 0, as_stream(stream), table, index, n, out);
  VAPC_CHECK_LAUNCH();
  return VAPC_OK;
}

extern "C" int vapc_point_distance(const double* x, const double* y, const double* z, int64_t n, const uint32_t* point2voxel,
                                   const double* cog3, int64_t nvoxels, double* distance, void* stream) {
  if (n < 0) return vapc_set_error(VAPC_E_INVALID, "vapc_point_distance: n < 0");
  if (n == 0) return VAPC_OK;
  VAPC_LAUNCH(point_distance_kernel, grid_for(n, 256, 4), 256, 0, as_stream(stream), x, y, z, n, point2voxel, cog3, nvoxels, distance);
  VAPC_CHECK_LAUNCH();
  return VAPC_OK;
}

// ---- host self-tests of the shared fp64 helpers (no GPU needed; used by the CPU test-suite) --------
extern "C" int vapc_selftest_eig3_host(const double* cov6_host, int64_t n, double* eig3_host) {
  for (int64_t i = 0; i < n; ++i) {
    const double* c = cov6_host + 6 * i;
    eig3_sym(c[0], c[1], c[2], c[3], c[4], c[5], eig3_host[3 * i], eig3_host[3 * i + 1], eig3_host[3 * i + 2]);
  }
  return VAPC_OK;
}
extern "C" int vapc_selftest_chan_host(const double* rows_host /*[n][10]: n, mean3, m2_6*/, int64_t n, double* out10_host) {
  double na = 0.0, ma[3] = {0, 0, 0}, Ma[6] = {0, 0, 0, 0, 0, 0};
  for (int64_t i = 0; i < n; ++i) {
    const double* r = rows_host + 10 * i;
    const double mb[3] = {r[1], r[2], r[3]};
    const double Mb[6] = {r[4], r[5], r[6], r[7], r[8], r[9]};
    chan_combine(na, ma, Ma, r[0], mb, Mb);
  }
  out10_host[0] = na;
  for (int k = 0; k < 3; ++k) out10_host[1 + k] = ma[k];
  for (int k = 0; k < 6; ++k) out10_host[4 + k] = Ma[k];
  return VAPC_OK;
}
extern "C" int vapc_selftest_mahalanobis_host(const double* cov6_host, const double* diff3_host, int64_t n, double* d_host, double* p_host) {
  for (int64_t i = 0; i < n; ++i) {
    double c[6];
    for (int k = 0; k < 6; ++k) c[k] = cov6_host[6 * i + k];
    c[0] += 1e-10; c[3] += 1e-10; c[5] += 1e-10;
    d_host[i] = quad_form_inv3(c, diff3_host[3 * i], diff3_host[3 * i + 1], diff3_host[3 * i + 2]);
    p_host[i] = chi2_df3_one_minus_cdf(d_host[i]);
  }
  return VAPC_OK;
}
