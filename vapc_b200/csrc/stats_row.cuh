// Per-voxel statistics of ONE voxel row from its (count, mean - c(v), M2) moments: density, centroid, std, covariance, eigenvalues
// (cyclic Jacobi on the symmetric 3x3), the 3 normalised eigenvalues and the 8 eigenvalue features.  Shared by voxel_stats_kernel
// (stats.cu: rows of a finished table) and the multi-GPU owner merge (segment_reduce.cu: rows formed on the fly from the own and the
// received partial rows, so that the merged moments need not be written and read back).
//
// Reference formulas: vapc.py:740 (density), 843-844 (centroid), 877-878 (std, ddof = 1), 977-1006 (covariance, NaN when n <= 1),
// 1055-1065 (eigenvalues of the ddof=1 covariance, sorted descending, NaN when n == 1), 1112-1145 (features).
#pragma once
#include "common.cuh"
#include "math3.cuh"

// Output rows of the axis-indexed tables.  The table's keys / moments may live in a permuted axis order (multi-GPU: the axis the
// chunks are separated along leads the key, vapc_b200/dist.py); the outputs are always x, y, z: internal axis i -> row axis[i] of
// cog / std, internal covariance component k -> component cov[k].
struct StatsAxes {
  int axis[3];
  int cov[6];
};
struct StatsOut {
  double *density, *cog3, *std3, *cov6, *eig3, *eign3, *feat8;
  int64_t nvox;  // column stride of the outputs
  StatsAxes ax;
  __host__ __device__ bool needs_m2() const { return std3 || cov6 || eig3 || eign3 || feat8; }
};
static inline bool stats_axes_from_order(const int32_t* axis_order_host, StatsAxes* ax) {
  *ax = StatsAxes{{0, 1, 2}, {0, 1, 2, 3, 4, 5}};
  if (!axis_order_host) return true;
  int seen = 0;
  for (int i = 0; i < 3; ++i) {
    if (axis_order_host[i] < 0 || axis_order_host[i] > 2) return false;
    ax->axis[i] = axis_order_host[i];
    seen |= 1 << axis_order_host[i];
  }
  if (seen != 7) return false;
  static const int pair[3][3] = {{0, 1, 2}, {1, 3, 4}, {2, 4, 5}};  // (a, b) -> xx xy xz yy yz zz
  int k = 0;
  for (int i = 0; i < 3; ++i)
    for (int j = i; j < 3; ++j) ax->cov[k++] = pair[ax->axis[i]][ax->axis[j]];
  return true;
}

// v = output row.  mean / M are only read when the outputs need them (cog3: mean; everything else: M).
__device__ __forceinline__ void emit_voxel_stats(unsigned long long key, unsigned n, const double (&mean)[3], const double (&M)[6], const Grid& g,
                                                 int64_t v, const StatsOut& o) {
  const int64_t nvox = o.nvox;
  const double nan = __longlong_as_double(0x7ff8000000000000LL);
  if (o.density) o.density[v] = __ddiv_rn((double)n, g.vs3);  // point_count / voxel_size ** 3, the cube as Python's pow gives it
  if (o.cog3) {
    long long vx, vy, vz;
    unpack_key(key, g, vx, vy, vz);
    o.cog3[o.ax.axis[0] * nvox + v] = shift_center(vx, g.ox, g.vs) + mean[0];
    o.cog3[o.ax.axis[1] * nvox + v] = shift_center(vy, g.oy, g.vs) + mean[1];
    o.cog3[o.ax.axis[2] * nvox + v] = shift_center(vz, g.oz, g.vs) + mean[2];
  }
  if (!o.needs_m2()) return;
  double c[6];
  if (n > 1) {
    const double d = 1.0 / (double)(n - 1);  // one reciprocal instead of six divisions (<= 1 ulp apart)
#pragma unroll
    for (int k = 0; k < 6; ++k) c[k] = M[k] * d;
  } else {
#pragma unroll
    for (int k = 0; k < 6; ++k) c[k] = nan;
  }
  if (o.cov6) {
#pragma unroll
    for (int k = 0; k < 6; ++k) o.cov6[(int64_t)o.ax.cov[k] * nvox + v] = c[k];
  }
  if (o.std3) {
    // fmax(NaN, 0) would hide the n == 1 NaN, so clamp only real values (rounding can leave -1e-33)
    o.std3[o.ax.axis[0] * nvox + v] = c[0] < 0.0 ? 0.0 : sqrt(c[0]);
    o.std3[o.ax.axis[1] * nvox + v] = c[3] < 0.0 ? 0.0 : sqrt(c[3]);
    o.std3[o.ax.axis[2] * nvox + v] = c[5] < 0.0 ? 0.0 : sqrt(c[5]);
  }
  if (!(o.eig3 || o.eign3 || o.feat8)) return;
  double l1 = nan, l2 = nan, l3 = nan;
  if (n > 1) {
    eig3_sym(c[0], c[1], c[2], c[3], c[4], c[5], l1, l2, l3);
    // a covariance is positive semi-definite: rounding noise below zero is clamped (the reference's
    // dgeev returns +-1e-18-sized noise of either sign there; see DESIGN.md, parity rule H7)
    l2 = fmax(l2, 0.0);
    l3 = fmax(l3, 0.0);
  }
  if (o.eig3) { o.eig3[v] = l1; o.eig3[nvox + v] = l2; o.eig3[2 * nvox + v] = l3; }
  if (!(o.eign3 || o.feat8)) return;
  const double total = l1 + l2 + l3;
  const double inv_total = 1.0 / total, inv_l1 = 1.0 / l1;  // 0 / 0 = NaN and x / 0 = inf behave the same through the reciprocal
  const double e1 = l1 * inv_total, e2 = l2 * inv_total, e3 = l3 * inv_total;
  if (o.eign3) { o.eign3[v] = e1; o.eign3[nvox + v] = e2; o.eign3[2 * nvox + v] = e3; }
  if (o.feat8) {
    double* f = o.feat8;
    f[v] = total;                                                                       // Sum_of_Eigenvalues
    f[nvox + v] = cbrt(l1 * l2 * l3);                                                   // Omnivariance ((.) ** (1/3), non-negative argument)
    f[2 * nvox + v] = -(e1 * safe_log(e1) + e2 * safe_log(e2) + e3 * safe_log(e3));     // Eigenentropy
    f[3 * nvox + v] = (l1 - l3) * inv_l1;                                               // Anisotropy
    f[4 * nvox + v] = (l2 - l3) * inv_l1;                                               // Planarity
    f[5 * nvox + v] = (l1 - l2) * inv_l1;                                               // Linearity
    f[6 * nvox + v] = e3;                                                               // Surface_Variation
    f[7 * nvox + v] = l3 * inv_l1;                                                      // Sphericity
  }
}
