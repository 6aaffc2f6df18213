// Device-wide exclusive scan of uint32 counters (tile head counts, digit counts).
// Reduce-then-scan in three launches; the arrays scanned here are N/1024-sized or smaller,
// so this is never on the bandwidth-critical path.
#include "common.cuh"

namespace {
constexpr int kThreads = 256;
constexpr int kItems = 16;
constexpr int kChunk = kThreads * kItems;  // 4096 elements per block

__global__ void __launch_bounds__(kThreads) scan_block_sums(const uint32_t* __restrict__ in, int64_t n,
                                                            unsigned long long* __restrict__ block_sums) {
  __shared__ unsigned long long warp_sum[kThreads / 32];
  const int64_t base = (int64_t)blockIdx.x * kChunk;
  unsigned long long s = 0;
#pragma unroll
  for (int k = 0; k < kItems; ++k) {
    const int64_t i = base + k * kThreads + threadIdx.x;
    if (i < n) s += in[i];
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if ((threadIdx.x & 31) == 0) warp_sum[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    unsigned long long t = 0;
    for (int w = 0; w < kThreads / 32; ++w) t += warp_sum[w];
    block_sums[blockIdx.x] = t;
  }
}

// One block turns block_sums into exclusive prefixes (sequential over chunks of 1024).
__global__ void __launch_bounds__(1024) scan_block_prefix(unsigned long long* __restrict__ block_sums, int nblocks,
                                                          unsigned long long* __restrict__ total) {
  __shared__ unsigned long long warp_tot[32];
  __shared__ unsigned long long carry_s;
  if (threadIdx.x == 0) carry_s = 0;
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int base = 0; base < nblocks; base += 1024) {
    const int i = base + threadIdx.x;
    const unsigned long long v = i < nblocks ? block_sums[i] : 0ull;
    unsigned long long inc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      unsigned long long t = __shfl_up_sync(0xffffffffu, inc, o);
      if (lane >= o) inc += t;
    }
    if (lane == 31) warp_tot[warp] = inc;
    __syncthreads();
    if (warp == 0) {
      unsigned long long w = warp_tot[lane];
      unsigned long long winc = w;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        unsigned long long t = __shfl_up_sync(0xffffffffu, winc, o);
        if (lane >= o) winc += t;
      }
      warp_tot[lane] = winc - w;
    }
    __syncthreads();
    const unsigned long long carry = carry_s;
    if (i < nblocks) block_sums[i] = carry + warp_tot[warp] + inc - v;
    __syncthreads();
    if (threadIdx.x == 1023) carry_s = carry + warp_tot[warp] + inc;
    __syncthreads();
  }
  if (threadIdx.x == 0 && total) *total = carry_s;
}

__global__ void __launch_bounds__(kThreads) scan_apply(const uint32_t* __restrict__ in, uint32_t* __restrict__ out,
                                                       int64_t n, const unsigned long long* __restrict__ block_prefix) {
  __shared__ unsigned warp_s[33];
  const int64_t base = (int64_t)blockIdx.x * kChunk;
  // blocked arrangement: thread t owns items [t*kItems, (t+1)*kItems)
  uint32_t v[kItems];
  unsigned local = 0;
#pragma unroll
  for (int k = 0; k < kItems; ++k) {
    const int64_t i = base + (int64_t)threadIdx.x * kItems + k;
    v[k] = i < n ? in[i] : 0u;
    local += v[k];
  }
  unsigned total;
  const unsigned excl = block_exclusive_scan(local, warp_s, &total);
  uint32_t run = (uint32_t)(block_prefix[blockIdx.x] + (unsigned long long)excl);
#pragma unroll
  for (int k = 0; k < kItems; ++k) {
    const int64_t i = base + (int64_t)threadIdx.x * kItems + k;
    if (i < n) out[i] = run;
    run += v[k];
  }
}
}  // namespace

size_t scan_u32_workspace_bytes(int64_t n) {
  const int64_t nb = (n + kChunk - 1) / kChunk;
  return (size_t)(nb > 0 ? nb : 1) * sizeof(unsigned long long);
}

cudaError_t scan_u32_exclusive(const uint32_t* in, uint32_t* out, int64_t n, unsigned long long* total, void* ws,
                               cudaStream_t s) {
  if (n <= 0) {
    if (total) return cudaMemsetAsync(total, 0, sizeof(unsigned long long), s);
    return cudaSuccess;
  }
  const int nb = (int)((n + kChunk - 1) / kChunk);
  unsigned long long* sums = static_cast<unsigned long long*>(ws);
  VAPC_LAUNCH(scan_block_sums, nb, kThreads, 0, s, in, n, sums);
  VAPC_LAUNCH(scan_block_prefix, 1, 1024, 0, s, sums, nb, total);
  VAPC_LAUNCH(scan_apply, nb, kThreads, 0, s, in, out, n, sums);
  return cudaGetLastError();
}
