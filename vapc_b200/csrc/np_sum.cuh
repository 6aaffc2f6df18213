// numpy's fp64 add.reduce over a 1-D slice, operation for operation (numpy/_core/src/umath/loops_utils.h.src, pairwise_sum):
//   fewer than 8 values: left to right from -0.0;  up to 128: eight interleaved accumulators r[j] += a[i + j], combined as
//   ((r0 + r1) + (r2 + r3)) + ((r4 + r5) + (r6 + r7)), then the tail left to right;  more: split at n / 2 rounded down to a multiple
//   of 8 and add the two halves' sums.  The split recursion runs on an explicit stack (depth <= log2(n / 128) + 1 < 48): device
//   recursion would need a stack-size guess.  f(p) = the p-th value; every addition is one rounding (no FMA contraction).
// The same text compiles for the host (no __CUDACC__: plain IEEE operations, build with -ffp-contract=off), which is how
// tests/test_cpu_abi.py pins this arithmetic against ndarray.sum / mean / var without a GPU.
#pragma once
#include <cstdint>
#ifdef __CUDACC__
#define VAPC_NP_FN __device__
#define vapc_np_add(a, b) __dadd_rn((a), (b))
#else
#define VAPC_NP_FN
static inline double vapc_np_add(double a, double b) { return a + b; }
#endif

template <typename F>
VAPC_NP_FN inline double np_sum_block(F f, int64_t a, int64_t n) {
  if (n < 8) {
    double r = -0.0;
    for (int64_t i = 0; i < n; ++i) r = vapc_np_add(r, f(a + i));
    return r;
  }
  double r[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) r[j] = f(a + j);
  int64_t i = 8;
  for (; i < n - (n % 8); i += 8) {
#pragma unroll
    for (int j = 0; j < 8; ++j) r[j] = vapc_np_add(r[j], f(a + i + j));
  }
  double res = vapc_np_add(vapc_np_add(vapc_np_add(r[0], r[1]), vapc_np_add(r[2], r[3])), vapc_np_add(vapc_np_add(r[4], r[5]), vapc_np_add(r[6], r[7])));
  for (; i < n; ++i) res = vapc_np_add(res, f(a + i));
  return res;
}
template <typename F>
VAPC_NP_FN inline double np_pairwise_sum(F f, int64_t a0, int64_t n0) {
  if (n0 <= 128) return np_sum_block(f, a0, n0);
  constexpr int kDepth = 48;
  int64_t sa[kDepth], sn[kDepth];
  double left[kDepth];
  unsigned char state[kDepth];  // 0: not entered, 1: left half returned, 2: right half returned
  int sp = 0;
  sa[0] = a0; sn[0] = n0; state[0] = 0;
  double ret = 0.0;
  while (sp >= 0) {
    if (state[sp] == 0) {
      if (sn[sp] <= 128) {
        ret = np_sum_block(f, sa[sp], sn[sp]);
        --sp;
      } else {
        int64_t n2 = sn[sp] / 2;
        n2 -= n2 % 8;
        state[sp] = 1;
        sa[sp + 1] = sa[sp]; sn[sp + 1] = n2; state[sp + 1] = 0;
        ++sp;
      }
    } else if (state[sp] == 1) {
      left[sp] = ret;
      int64_t n2 = sn[sp] / 2;
      n2 -= n2 % 8;
      state[sp] = 2;
      sa[sp + 1] = sa[sp] + n2; sn[sp + 1] = sn[sp] - n2; state[sp + 1] = 0;
      ++sp;
    } else {
      ret = vapc_np_add(left[sp], ret);
      --sp;
    }
  }
  return ret;
}
