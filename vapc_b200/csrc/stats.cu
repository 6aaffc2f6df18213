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

namespace {
constexpr int kThreads = 128;
// Output rows of the axis-indexed tables.  The table's keys / moments may live in a permuted axis order (multi-GPU: the axis the
// chunks are separated along leads the key, vapc_b200/dist.py); the outputs are always x, y, z: internal axis i -> row axis[i] of
// cog / std, internal covariance component k -> component cov[k].
struct StatsAxes {
  int axis[3];
  int cov[6];
};

template <typename KeyT>
__global__ void __launch_bounds__(kThreads) voxel_stats_kernel(const KeyT* __restrict__ keys, const uint32_t* __restrict__ count,
                                                               const double* __restrict__ mean3, const double* __restrict__ m2,
                                                               int64_t nvox, Grid g, double* __restrict__ density,
                                                               double* __restrict__ cog3, double* __restrict__ std3,
                                                               double* __restrict__ cov6, double* __restrict__ eig3,
                                                               double* __restrict__ eign3, double* __restrict__ feat8, StatsAxes ax) {
  const int64_t v = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (v >= nvox) return;
  const double nan = __longlong_as_double(0x7ff8000000000000LL);
  const unsigned n = count[v];
  if (density) density[v] = (double)n / (g.vs * g.vs * g.vs);  // point_count / voxel_size**3
  if (cog3) {
    long long vx, vy, vz;
    unpack_key((unsigned long long)keys[v], g, vx, vy, vz);
    cog3[ax.axis[0] * nvox + v] = shift_center(vx, g.ox, g.vs) + mean3[v];
    cog3[ax.axis[1] * nvox + v] = shift_center(vy, g.oy, g.vs) + mean3[nvox + v];
    cog3[ax.axis[2] * nvox + v] = shift_center(vz, g.oz, g.vs) + mean3[2 * nvox + v];
  }
  if (!(std3 || cov6 || eig3 || eign3 || feat8)) return;
  double c[6];
  if (n > 1) {
    const double d = 1.0 / (double)(n - 1);  // one reciprocal instead of six divisions (<= 1 ulp apart)
#pragma unroll
    for (int k = 0; k < 6; ++k) c[k] = m2[(int64_t)k * nvox + v] * d;
  } else {
#pragma unroll
    for (int k = 0; k < 6; ++k) c[k] = nan;
  }
  if (cov6) {
#pragma unroll
    for (int k = 0; k < 6; ++k) cov6[(int64_t)ax.cov[k] * nvox + v] = c[k];
  }
  if (std3) {
    // fmax(NaN, 0) would hide the n == 1 NaN, so clamp only real values (rounding can leave -1e-33)
    std3[ax.axis[0] * nvox + v] = c[0] < 0.0 ? 0.0 : sqrt(c[0]);
    std3[ax.axis[1] * nvox + v] = c[3] < 0.0 ? 0.0 : sqrt(c[3]);
    std3[ax.axis[2] * nvox + v] = c[5] < 0.0 ? 0.0 : sqrt(c[5]);
  }
  if (!(eig3 || eign3 || feat8)) return;
  double l1 = nan, l2 = nan, l3 = nan;
  if (n > 1) {
    eig3_sym(c[0], c[1], c[2], c[3], c[4], c[5], l1, l2, l3);
    // a covariance is positive semi-definite: rounding noise below zero is clamped (the reference's
    // dgeev returns +-1e-18-sized noise of either sign there; see DESIGN.md, parity rule H7)
    l2 = fmax(l2, 0.0);
    l3 = fmax(l3, 0.0);
  }
  if (eig3) { eig3[v] = l1; eig3[nvox + v] = l2; eig3[2 * nvox + v] = l3; }
  if (!(eign3 || feat8)) return;
  const double total = l1 + l2 + l3;
  const double inv_total = 1.0 / total, inv_l1 = 1.0 / l1;  // 0 / 0 = NaN and x / 0 = inf behave the same through the reciprocal
  const double e1 = l1 * inv_total, e2 = l2 * inv_total, e3 = l3 * inv_total;
  if (eign3) { eign3[v] = e1; eign3[nvox + v] = e2; eign3[2 * nvox + v] = e3; }
  if (feat8) {
    feat8[v] = total;                                                                       // Sum_of_Eigenvalues
    feat8[nvox + v] = cbrt(l1 * l2 * l3);                                                   // Omnivariance ((.) ** (1/3), non-negative argument)
    feat8[2 * nvox + v] = -(e1 * safe_log(e1) + e2 * safe_log(e2) + e3 * safe_log(e3));     // Eigenentropy
    feat8[3 * nvox + v] = (l1 - l3) * inv_l1;                                               // Anisotropy
    feat8[4 * nvox + v] = (l2 - l3) * inv_l1;                                               // Planarity
    feat8[5 * nvox + v] = (l1 - l2) * inv_l1;                                               // Linearity
    feat8[6 * nvox + v] = e3;                                                               // Surface_Variation
    feat8[7 * nvox + v] = l3 * inv_l1;                                                      // Sphericity
  }
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
  StatsAxes ax = {{0, 1, 2}, {0, 1, 2, 3, 4, 5}};
  if (axis_order_host) {
    int seen = 0;
    for (int i = 0; i < 3; ++i) {
      if (axis_order_host[i] < 0 || axis_order_host[i] > 2) return vapc_set_error(VAPC_E_INVALID, "vapc_voxel_stats_axes: axis_order must be a permutation of 0, 1, 2");
      ax.axis[i] = axis_order_host[i];
      seen |= 1 << axis_order_host[i];
    }
    if (seen != 7) return vapc_set_error(VAPC_E_INVALID, "vapc_voxel_stats_axes: axis_order must be a permutation of 0, 1, 2");
    static const int pair[3][3] = {{0, 1, 2}, {1, 3, 4}, {2, 4, 5}};  // (a, b) -> xx xy xz yy yz zz
    int k = 0;
    for (int i = 0; i < 3; ++i)
      for (int j = i; j < 3; ++j) ax.cov[k++] = pair[ax.axis[i]][ax.axis[j]];
  }
  if (nvoxels == 0) return VAPC_OK;
  if (!voxel_keys || !count || !mean3 || !m2) return vapc_set_error(VAPC_E_INVALID, "vapc_voxel_stats: null input");
  const Grid g = to_device_grid(grid_host);
  const unsigned blocks = (unsigned)((nvoxels + kThreads - 1) / kThreads);
  if (grid_host->key_bytes == 4)
    VAPC_LAUNCH(voxel_stats_kernel<uint32_t>, blocks, kThreads, 0, as_stream(stream), static_cast<const uint32_t*>(voxel_keys), count, mean3, m2,
                                                                             nvoxels, g, density, cog3, std3, cov6, eig3, eign3, feat8, ax);
  else
    VAPC_LAUNCH(voxel_stats_kernel<unsigned long long>, blocks, kThreads, 0, as_stream(stream), 
        static_cast<const unsigned long long*>(voxel_keys), count, mean3, m2, nvoxels, g, density, cog3, std3, cov6, eig3, eign3, feat8, ax);
  VAPC_CHECK_LAUNCH();
  return VAPC_OK;
}

extern "C" int vapc_voxel_stats(const void* voxel_keys, const uint32_t* count, const double* mean3, const double* m2,
                                int64_t nvoxels, const vapc_grid_t* grid_host, double* density, double* cog3, double* std3,
                                double* cov6, double* eig3, double* eign3, double* feat8, void* stream) {
  return vapc_voxel_stats_axes(voxel_keys, count, mean3, m2, nvoxels, grid_host, nullptr, density, cog3, std3, cov6, eig3, eign3, feat8, stream);
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
