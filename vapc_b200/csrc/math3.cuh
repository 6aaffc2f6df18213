// Small fp64 3x3 helpers shared by device kernels and the host self-test entry points
// (vapc_selftest_*_host): compiled for both sides so the arithmetic can be exercised without a GPU.
#pragma once
#include <math.h>

#ifdef __CUDACC__
#define VAPC_HD __host__ __device__ __forceinline__
#else
#define VAPC_HD inline
#endif

VAPC_HD double vapc_rsqrt(double v) {
#ifdef __CUDA_ARCH__
  return rsqrt(v);
#else
  return 1.0 / sqrt(v);
#endif
}

// One Jacobi rotation annihilating a_pq of a symmetric 3x3 held in scalars.
// (app, aqq, apq) is the 2x2 pivot block, (arp, arq) the remaining off-diagonal pair.
// The rotation angle (|angle| <= pi/4) comes from its double angle, with two reciprocal square roots and no division or square
// root (an fp64 division is ~30 instructions on the GPU, a square root ~20, rsqrt ~10, and the kernel running this ~8 times per
// voxel is bound by issue slots):  with h = aqq - app and rho = 1 / sqrt(h^2 + 4 apq^2)
//   cos 2a = |h| rho,  sin 2a = sgn(h) 2 apq rho,  cos^2 a = (1 + cos 2a) / 2,  cos a = cos^2 a * rsqrt(cos^2 a),
//   sin a = sin 2a / (2 cos a),  tan a = sin a / cos a.
// The matrix is pre-scaled to max |entry| = 1 and the caller only rotates pivots above its convergence threshold, so
// h^2 + 4 apq^2 neither overflows nor underflows.
VAPC_HD void jacobi_rotate(double& app, double& aqq, double& apq, double& arp, double& arq) {
  const double h = aqq - app;
  const double rho = vapc_rsqrt(fma(h, h, 4.0 * apq * apq));
  const double c2 = fma(0.5 * fabs(h), rho, 0.5);  // cos^2 in [1/2, 1]
  const double ic = vapc_rsqrt(c2);                // 1 / cos
  const double c = c2 * ic;
  const double s = copysign(apq * rho, apq * copysign(1.0, h)) * ic;
  const double hh = (s * ic) * apq;                // tan * apq
  app -= hh;
  aqq += hh;
  apq = 0.0;
  const double gp = arp, hq = arq;
  arp = fma(c, gp, -s * hq);
  arq = fma(s, gp, c * hq);
}

// Eigenvalues of the symmetric matrix [[a00 a01 a02][a01 a11 a12][a02 a12 a22]], descending.
// CLASSICAL Jacobi on the matrix scaled to max |entry| = 1: every step annihilates the LARGEST off-diagonal element.  The pivot
// is brought to position (0, 1) by relabelling the axes (conditional swaps: selects on the integer pipe — eigenvalues do not care
// about the labelling), so one rotation routine serves all three pivots without divergent branches.  The kernel using this is bound
// by the FP64 pipe; a warp pays for its slowest lane: 7.7 rotations per warp on average against 12-15 rotation slots of the cyclic
// sweeps this replaces (measured on covariances of 2-13 points, /tmp prototype quoted in DESIGN.md).  Stop at an off-diagonal mass of
// 1e-11: |d lambda| <= |E| bounds the error by 1e-11 of the scale in the worst (clustered) case, and Jacobi converges
// quadratically, so the typical error is ~1e-22; the parity rule is 1e-9 of the largest eigenvalue (H7).
VAPC_HD void eig3_sym(double a00, double a01, double a02, double a11, double a12, double a22, double& l1,
                                         double& l2, double& l3) {
  const double scale = fmax(fmax(fabs(a00), fabs(a11)), fmax(fabs(a22), fmax(fabs(a01), fmax(fabs(a02), fabs(a12)))));
  if (scale == 0.0) { l1 = l2 = l3 = 0.0; return; }
  const double inv = 1.0 / scale;
  a00 *= inv; a01 *= inv; a02 *= inv; a11 *= inv; a12 *= inv; a22 *= inv;
  for (int it = 0; it < 64; ++it) {
    const double f01 = fabs(a01), f02 = fabs(a02), f12 = fabs(a12);
    if (f01 + f02 + f12 <= 1e-11) break;
    const bool p02 = f02 >= f01 && f02 >= f12;  // pivot (0, 2): swap the labels of axes 1 and 2
    const bool p12 = !p02 && f12 > f01;         // pivot (1, 2): swap the labels of axes 0 and 2
    double t;
    t = p02 ? a22 : a11; a22 = p02 ? a11 : a22; a11 = t;
    t = p02 ? a02 : a01; a02 = p02 ? a01 : a02; a01 = t;
    t = p12 ? a22 : a00; a22 = p12 ? a00 : a22; a00 = t;
    t = p12 ? a12 : a01; a12 = p12 ? a01 : a12; a01 = t;
    jacobi_rotate(a00, a11, a01, a02, a12);  // (p, q) = (0, 1), r = 2
  }
  double x = a00 * scale, y = a11 * scale, z = a22 * scale;
  double t;
  if (x < y) { t = x; x = y; y = t; }
  if (y < z) { t = y; y = z; z = t; }
  if (x < y) { t = x; x = y; y = t; }
  l1 = x; l2 = y; l3 = z;
}

VAPC_HD double safe_log(double v) { return v > 0.0 ? log(v) : 0.0; }  // vapc.py:1104-1106


// ---- Chan / Welford pairwise combination of (n, mean, M2) -----------------------------------------
// m: mean of (p - c(voxel)); M: xx xy xz yy yz zz central second moments.
VAPC_HD void chan_combine(double& na, double (&ma)[3], double (&Ma)[6], double nb, const double (&mb)[3], const double (&Mb)[6]) {
  if (nb == 0.0) return;
  if (na == 0.0) {
    na = nb;
    for (int k = 0; k < 3; ++k) ma[k] = mb[k];
    for (int k = 0; k < 6; ++k) Ma[k] = Mb[k];
    return;
  }
  const double n = na + nb;
  const double f = nb / n, w = na * f;
  const double dx = mb[0] - ma[0], dy = mb[1] - ma[1], dz = mb[2] - ma[2];
  ma[0] = fma(dx, f, ma[0]); ma[1] = fma(dy, f, ma[1]); ma[2] = fma(dz, f, ma[2]);
  Ma[0] += fma(dx * dx, w, Mb[0]); Ma[1] += fma(dx * dy, w, Mb[1]); Ma[2] += fma(dx * dz, w, Mb[2]);
  Ma[3] += fma(dy * dy, w, Mb[3]); Ma[4] += fma(dy * dz, w, Mb[4]); Ma[5] += fma(dz * dz, w, Mb[5]);
  na = n;
}

// ---- q = d' inv(C) d for a symmetric 3x3 C = [[c0 c1 c2][c1 c3 c4][c2 c4 c5]] -------------------------
// (np.linalg.inv + einsum in the reference, bi_vapc.py:112-117).  Solved as C y = d by Gaussian elimination
// with partial pivoting — the same backward-stable class as LAPACK's getrf the reference runs; the
// adjugate / determinant formula loses cond(C)^2 and is useless for rank-deficient covariances + 1e-10 I.
VAPC_HD double quad_form_inv3(const double (&c)[6], double dx, double dy, double dz) {
  double r0[4] = {c[0], c[1], c[2], dx};
  double r1[4] = {c[1], c[3], c[4], dy};
  double r2[4] = {c[2], c[4], c[5], dz};
#define VAPC_SWAP_ROWS(a, b)                                  \
  do {                                                        \
    for (int k_ = 0; k_ < 4; ++k_) {                          \
      const double t_ = a[k_]; a[k_] = b[k_]; b[k_] = t_;     \
    }                                                         \
  } while (0)
  // column 0 pivot
  if (fabs(r1[0]) > fabs(r0[0])) VAPC_SWAP_ROWS(r0, r1);
  if (fabs(r2[0]) > fabs(r0[0])) VAPC_SWAP_ROWS(r0, r2);
  const double m1 = r1[0] / r0[0], m2 = r2[0] / r0[0];
  for (int k = 1; k < 4; ++k) { r1[k] -= m1 * r0[k]; r2[k] -= m2 * r0[k]; }
  // column 1 pivot
  if (fabs(r2[1]) > fabs(r1[1])) VAPC_SWAP_ROWS(r1, r2);
  const double m3 = r2[1] / r1[1];
  for (int k = 2; k < 4; ++k) r2[k] -= m3 * r1[k];
#undef VAPC_SWAP_ROWS
  const double y2 = r2[3] / r2[2];
  const double y1 = (r1[3] - r1[2] * y2) / r1[1];
  const double y0 = (r0[3] - r0[1] * y1 - r0[2] * y2) / r0[0];
  return dx * y0 + dy * y1 + dz * y2;
}

// p = 1 - chi2.cdf(d, df = 3) (bi_vapc.py:118-119).  Closed form of the upper tail for 3 degrees of freedom:
//   q = erfc(sqrt(d / 2)) + sqrt(2 d / pi) * exp(-d / 2)
// scipy forms cdf = 1 - q (cephes igam: the complemented continued fraction for large arguments) and the reference then takes
// 1 - cdf; the two roundings are reproduced here ON PURPOSE: they make p exactly 0 once q <= 2^-54 (d >= 78.80...), and
// p == 0 is what the reference turns into change_type 2 (bi_vapc.py:135).  Summing the two tail terms (instead of
// 1 - (erf - ...)) puts that threshold at the same d as scipy's; measured on the host against scipy: |p - p_ref| <= 4.3e-15
// absolute and <= 1e-9 relative over d in (1e-8, 100), identical zero sets on 4e5 samples across the threshold.
VAPC_HD double chi2_df3_one_minus_cdf(double d) {
  if (d != d) return d;
  if (d <= 0.0) return 1.0;
  const double q = erfc(sqrt(0.5 * d)) + sqrt(d * 0.63661977236758134308 /* 2/pi */) * exp(-0.5 * d);
  const double cdf = fmax(1.0 - q, 0.0);  // q can round to 1 + 1e-16 for tiny d; a cdf is not negative
  return 1.0 - cdf;
}
