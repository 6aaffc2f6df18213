// Two-epoch voxel join + centroid delta + Mahalanobis / chi-square change classification.
// Reference: bi_vapc.py:78-84 (inner join on the voxel triple), :25-36 (anti-joins), :86-88 / :198-207
// (Euclidean centroid distance), :92-138 (Mahalanobis test).  Everything here is per-voxel work
// (V-sized, not N-sized).
#include "common.cuh"
#include "math3.cuh"

namespace {
constexpr int kThreads = 256;

__device__ __forceinline__ uint32_t lower_bound_match(const unsigned long long* __restrict__ keys, int64_t n, unsigned long long k) {
  int64_t lo = 0, hi = n;
  while (lo < hi) {
    const int64_t mid = (lo + hi) >> 1;
    if (__ldg(keys + mid) < k) lo = mid + 1; else hi = mid;
  }
  return (lo < n && __ldg(keys + lo) == k) ? (uint32_t)lo : 0xffffffffu;
}

__global__ void __launch_bounds__(kThreads) join_kernel(const unsigned long long* __restrict__ a, int64_t na,
                                                        const unsigned long long* __restrict__ b, int64_t nb,
                                                        uint32_t* __restrict__ match_a, uint32_t* __restrict__ match_b) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < na) match_a[i] = lower_bound_match(b, nb, a[i]);
  if (i < nb) match_b[i] = lower_bound_match(a, na, b[i]);
}

__global__ void __launch_bounds__(kThreads) mahalanobis_kernel(const double* __restrict__ cog1, const double* __restrict__ cov1, int64_t v1,
                                                               const double* __restrict__ cog2, const double* __restrict__ cov2, int64_t v2,
                                                               const uint32_t* __restrict__ i1, const uint32_t* __restrict__ i2,
                                                               int64_t npairs, double alpha, double* __restrict__ distance,
                                                               double* __restrict__ d1_out, double* __restrict__ d2_out,
                                                               double* __restrict__ p_value, int32_t* __restrict__ significance,
                                                               int32_t* __restrict__ change_type) {
  const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= npairs) return;
  const int64_t a = i1[r], b = i2[r];
  const double dx = cog1[a] - cog2[b], dy = cog1[v1 + a] - cog2[v2 + b], dz = cog1[2 * v1 + a] - cog2[2 * v2 + b];
  if (distance) distance[r] = __dsqrt_rn(__dadd_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)), __dmul_rn(dz, dz)));
  if (!(d1_out || d2_out || p_value || significance || change_type)) return;
  double cx[6], cy[6];
#pragma unroll
  for (int k = 0; k < 6; ++k) { cx[k] = cov1[(int64_t)k * v1 + a]; cy[k] = cov2[(int64_t)k * v2 + b]; }
  const double eps = 1e-10;  // bi_vapc.py:108-110
  cx[0] += eps; cx[3] += eps; cx[5] += eps;
  cy[0] += eps; cy[3] += eps; cy[5] += eps;
  const double d1 = quad_form_inv3(cy, dx, dy, dz);  // diff' inv(cov epoch 2) diff   (bi_vapc.py:116)
  const double d2 = quad_form_inv3(cx, dx, dy, dz);  // diff' inv(cov epoch 1) diff   (bi_vapc.py:117)
  const double p1 = chi2_df3_one_minus_cdf(d1), p2 = chi2_df3_one_minus_cdf(d2);
  const bool sig = (p1 < alpha) || (p2 < alpha);
  const double p = (p1 < alpha) ? p1 : p2;
  if (d1_out) d1_out[r] = d1;
  if (d2_out) d2_out[r] = d2;
  if (p_value) p_value[r] = p;
  if (significance) significance[r] = sig ? 1 : 0;
  if (change_type) change_type[r] = (p == 0.0) ? 2 : (sig ? 1 : 0);  // bi_vapc.py:132-135
}
}  // namespace

extern "C" int vapc_join_sorted_keys(const uint64_t* keys1, int64_t n1, const uint64_t* keys2, int64_t n2, uint32_t* match1,
                                     uint32_t* match2, void* stream) {
  if (n1 < 0 || n2 < 0) return vapc_set_error(VAPC_E_INVALID, "vapc_join_sorted_keys: negative size");
  const int64_t n = n1 > n2 ? n1 : n2;
  if (n == 0) return VAPC_OK;
  VAPC_LAUNCH(join_kernel, (unsigned)((n + kThreads - 1) / kThreads), kThreads, 0, as_stream(stream), 
      reinterpret_cast<const unsigned long long*>(keys1), n1, reinterpret_cast<const unsigned long long*>(keys2), n2, match1, match2);
  VAPC_CHECK_LAUNCH();
  return VAPC_OK;
}

extern "C" int vapc_mahalanobis(const double* cog3_1, const double* cov6_1, int64_t v1, const double* cog3_2, const double* cov6_2,
                                int64_t v2, const uint32_t* i1, const uint32_t* i2, int64_t npairs, double alpha, double* distance,
                                double* d1, double* d2, double* p_value, int32_t* significance, int32_t* change_type, void* stream) {
  if (npairs < 0) return vapc_set_error(VAPC_E_INVALID, "vapc_mahalanobis: negative size");
  if (npairs == 0) return VAPC_OK;
  if (!cog3_1 || !cog3_2 || !i1 || !i2) return vapc_set_error(VAPC_E_INVALID, "vapc_mahalanobis: null input");
  if ((d1 || d2 || p_value || significance || change_type) && (!cov6_1 || !cov6_2))
    return vapc_set_error(VAPC_E_INVALID, "vapc_mahalanobis: covariance tables required");
  VAPC_LAUNCH(mahalanobis_kernel, (unsigned)((npairs + kThreads - 1) / kThreads), kThreads, 0, as_stream(stream), 
      cog3_1, cov6_1, v1, cog3_2, cov6_2, v2, i1, i2, npairs, alpha, distance, d1, d2, p_value, significance, change_type);
  VAPC_CHECK_LAUNCH();
  return VAPC_OK;
}
