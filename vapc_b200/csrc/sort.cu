// In-house stable LSD radix sort of (key, point index) pairs — 8-bit digits.
//
// Per pass (tile = 4096 consecutive elements, one CTA of 256 threads per tile):
//   radix_hist     per-tile digit histogram              reads keys                 K B/elt
//   colscan_*      column-wise exclusive scan of the [ntiles][256] counters (tiny)
//   radix_scatter  rank (ballot-built peer masks + per-warp smem digit counters), reorder the tile in
//                  shared memory, write each digit's run coalesced   reads K+4, writes K+4 B/elt
// Peer masks ("which lanes hold my digit") are built from 8 vote.ballot's, one per digit bit.  The
// match.any instruction would do the same in one line, but it runs on the SM's single ADU pipe at a
// cost proportional to the number of distinct values in the warp (~60 cycles for random 8-bit digits;
// ncu showed both kernels 99 % ADU-bound with it, profiles/r1_ncu_summary.md).
// Ranking keeps equal digits in input order, so every pass — and therefore the whole sort — is
// stable: points of one voxel stay in original index order (tie-break "lowest original index"
// and "first point of a voxel" fall out of that; see include/vapc_b200.h).
#include "common.cuh"

namespace {
constexpr int kRadixBits = 8;
constexpr int kRadix = 1 << kRadixBits;
constexpr int kThreads = 256;  // == kRadix: thread t owns digit t in the per-digit steps
constexpr int kWarps = kThreads / 32;
constexpr int kItems = 16;
constexpr int kTile = kThreads * kItems;  // 4096
constexpr int kRowsPerChunk = 32;         // tiles per column-scan chunk

template <typename KeyT>
__device__ __forceinline__ unsigned digit_of(KeyT k, int shift, unsigned mask) {
  return (unsigned)(k >> shift) & mask;
}

// lanes of the warp holding the same 8-bit digit as this lane (invalid lanes match nobody valid)
__device__ __forceinline__ unsigned match_digit(unsigned d, bool valid) {
  unsigned peers = __ballot_sync(0xffffffffu, valid);
#pragma unroll
  for (int b = 0; b < kRadixBits; ++b) {
    const bool bit = (d >> b) & 1u;
    const unsigned m = __ballot_sync(0xffffffffu, bit);
    peers &= bit ? m : ~m;
  }
  return peers;
}

// ---- per-tile digit histogram ---------------------------------------------------------------
// Per-warp private counters updated by one leader lane per distinct digit (plain read-modify-write:
// leaders of different digits touch different words), then summed over the warps.
template <typename KeyT>
__global__ void __launch_bounds__(kThreads) radix_hist(const KeyT* __restrict__ keys, int64_t n, int shift, unsigned mask,
                                                       uint32_t* __restrict__ tile_hist) {
  __shared__ unsigned cnt[kWarps][kRadix];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int w = 0; w < kWarps; ++w) cnt[w][threadIdx.x] = 0;
  __syncthreads();
  const int64_t base = (int64_t)blockIdx.x * kTile;
  KeyT key[kItems];
#pragma unroll
  for (int k = 0; k < kItems; ++k) {
    const int64_t i = base + k * kThreads + threadIdx.x;
    key[k] = i < n ? keys[i] : (KeyT)0;
  }
  unsigned* wc = cnt[warp];
#pragma unroll
  for (int k = 0; k < kItems; ++k) {
    const bool valid = (base + k * kThreads + threadIdx.x) < n;
    const unsigned d = digit_of(key[k], shift, mask);
    const unsigned peers = match_digit(d, valid);
    if (valid && lane == (__ffs(peers) - 1)) wc[d] += (unsigned)__popc(peers);
    __syncwarp();
  }
  __syncthreads();
  unsigned t = 0;
#pragma unroll
  for (int w = 0; w < kWarps; ++w) t += cnt[w][threadIdx.x];
  tile_hist[(int64_t)blockIdx.x * kRadix + threadIdx.x] = t;
}

// ---- column scan: off[t][d] = SUM_{d'<d} total[d'] + SUM_{t'<t} hist[t'][d] ---------------------
// Chunks of kRowsPerChunk tiles; every thread owns one digit column.  All loads of a chunk are issued
// before the dependent running sum, so the passes are bandwidth- not latency-bound.
__global__ void __launch_bounds__(kThreads) colscan_partial(const uint32_t* __restrict__ hist, int ntiles,
                                                            uint32_t* __restrict__ chunk_sum) {
  const int r0 = blockIdx.x * kRowsPerChunk;
  uint32_t v[kRowsPerChunk];
#pragma unroll
  for (int r = 0; r < kRowsPerChunk; ++r) v[r] = (r0 + r) < ntiles ? hist[(int64_t)(r0 + r) * kRadix + threadIdx.x] : 0u;
  uint32_t s = 0;
#pragma unroll
  for (int r = 0; r < kRowsPerChunk; ++r) s += v[r];
  chunk_sum[(int64_t)blockIdx.x * kRadix + threadIdx.x] = s;
}
__global__ void __launch_bounds__(kThreads) colscan_chunks(const uint32_t* __restrict__ chunk_sum, uint32_t* __restrict__ chunk_base,
                                                           int nchunks) {
  __shared__ unsigned warp_s[33];
  uint32_t total_d = 0;
#pragma unroll 8
  for (int c = 0; c < nchunks; ++c) total_d += chunk_sum[(int64_t)c * kRadix + threadIdx.x];
  unsigned total;
  uint32_t run = block_exclusive_scan(total_d, warp_s, &total);  // first slot of this digit
#pragma unroll 8
  for (int c = 0; c < nchunks; ++c) {
    const uint32_t v = chunk_sum[(int64_t)c * kRadix + threadIdx.x];
    chunk_base[(int64_t)c * kRadix + threadIdx.x] = run;
    run += v;
  }
}
__global__ void __launch_bounds__(kThreads) colscan_apply(uint32_t* __restrict__ hist, int ntiles,
                                                          const uint32_t* __restrict__ chunk_base) {
  const int r0 = blockIdx.x * kRowsPerChunk;
  uint32_t v[kRowsPerChunk];
#pragma unroll
  for (int r = 0; r < kRowsPerChunk; ++r) v[r] = (r0 + r) < ntiles ? hist[(int64_t)(r0 + r) * kRadix + threadIdx.x] : 0u;
  uint32_t run = chunk_base[(int64_t)blockIdx.x * kRadix + threadIdx.x];
#pragma unroll
  for (int r = 0; r < kRowsPerChunk; ++r) {
    if ((r0 + r) < ntiles) hist[(int64_t)(r0 + r) * kRadix + threadIdx.x] = run;
    run += v[r];
  }
}

// ---- rank + scatter ---------------------------------------------------------------------------
template <typename KeyT>
struct ScatterSmem {
  unsigned warp_cnt[kWarps][kRadix];  // per-warp digit counters -> per-warp exclusive prefixes
  unsigned bin_start[kRadix];         // tile-local first slot of each digit
  unsigned g_off[kRadix];             // global first slot of this tile's run of each digit
  unsigned warp_s[33];
  KeyT keys[kTile];
  uint32_t idx[kTile];
};

template <typename KeyT, bool IOTA>
__global__ void __launch_bounds__(kThreads, 3) radix_scatter(const KeyT* __restrict__ keys_in, const uint32_t* __restrict__ idx_in,
                                                          KeyT* __restrict__ keys_out, uint32_t* __restrict__ idx_out, int64_t n,
                                                          int shift, unsigned mask, const uint32_t* __restrict__ tile_off) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  ScatterSmem<KeyT>& sm = *reinterpret_cast<ScatterSmem<KeyT>*>(smem_raw);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int64_t tile_base = (int64_t)blockIdx.x * kTile;

#pragma unroll
  for (int w = 0; w < kWarps; ++w) sm.warp_cnt[w][tid] = 0;
  sm.g_off[tid] = tile_off[(int64_t)blockIdx.x * kRadix + tid];

  // warp-striped load: warp w owns elements [w*512, (w+1)*512) of the tile, lane l takes
  // element k*32 + l in step k, so (warp, k, lane) order is input order.
  KeyT key[kItems];
  uint32_t idx[kItems];
  const int64_t wbase = tile_base + (int64_t)warp * (32 * kItems) + lane;
#pragma unroll
  for (int k = 0; k < kItems; ++k) {
    const int64_t i = wbase + k * 32;
    const bool valid = i < n;
    key[k] = valid ? keys_in[i] : (KeyT)0;
    idx[k] = IOTA ? (uint32_t)i : (valid ? idx_in[i] : 0u);
  }
  __syncthreads();

  // rank inside the warp: lanes holding the same digit find each other through the ballot-built peer
  // mask; the per-warp counter carries the count from earlier steps.
  unsigned rank[kItems];
  const unsigned lt = (1u << lane) - 1u;
  unsigned* wc = sm.warp_cnt[warp];
#pragma unroll
  for (int k = 0; k < kItems; ++k) {
    const bool valid = (wbase + k * 32) < n;
    const unsigned d = digit_of(key[k], shift, mask);
    const unsigned peers = match_digit(d, valid);
    unsigned prev = 0;
    if (valid) prev = wc[d];
    __syncwarp();
    if (valid && lane == (__ffs(peers) - 1)) wc[d] = prev + (unsigned)__popc(peers);
    __syncwarp();
    rank[k] = prev + (unsigned)__popc(peers & lt);
  }
  __syncthreads();

  // per digit (thread tid == digit): exclusive prefix over warps, then over digits
  unsigned run = 0;
#pragma unroll
  for (int w = 0; w < kWarps; ++w) {
    const unsigned t = sm.warp_cnt[w][tid];
    sm.warp_cnt[w][tid] = run;
    run += t;
  }
  unsigned total;
  sm.bin_start[tid] = block_exclusive_scan(run, sm.warp_s, &total);
  __syncthreads();

  // reorder inside shared memory
#pragma unroll
  for (int k = 0; k < kItems; ++k) {
    if ((wbase + k * 32) < n) {
      const unsigned d = digit_of(key[k], shift, mask);
      const unsigned pos = sm.bin_start[d] + wc[d] + rank[k];
      sm.keys[pos] = key[k];
      sm.idx[pos] = idx[k];
    }
  }
  __syncthreads();

  // each digit's run goes out contiguously
  const int tile_n = (int)min((int64_t)kTile, n - tile_base);
#pragma unroll 4
  for (int p = tid; p < tile_n; p += kThreads) {
    const KeyT k = sm.keys[p];
    const unsigned d = digit_of(k, shift, mask);
    const unsigned dst = sm.g_off[d] + ((unsigned)p - sm.bin_start[d]);
    keys_out[dst] = k;
    idx_out[dst] = sm.idx[p];
  }
}

__global__ void iota_kernel(uint32_t* __restrict__ out, int64_t n) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) out[i] = (uint32_t)i;
}

inline int64_t num_tiles(int64_t n) { return (n + kTile - 1) / kTile; }
inline int64_t num_chunks(int64_t ntiles) { return (ntiles + kRowsPerChunk - 1) / kRowsPerChunk; }

template <typename KeyT>
int sort_impl(KeyT* keys, KeyT* keys_alt, KeyT* keys_alt2, uint32_t* idx, uint32_t* idx_alt, int idx_is_iota, int64_t n, int end_bit,
              void* ws, int32_t* result_location, cudaStream_t s) {
  const int passes = (end_bit + kRadixBits - 1) / kRadixBits;
  const int64_t ntiles = num_tiles(n);
  const int64_t nchunks = num_chunks(ntiles);
  uint32_t* tile_hist = static_cast<uint32_t*>(ws);
  uint32_t* chunk_sum = tile_hist + ntiles * kRadix;
  uint32_t* chunk_base = chunk_sum + nchunks * kRadix;
  const size_t smem = sizeof(ScatterSmem<KeyT>);
  VAPC_RETURN_IF_CUDA(cudaFuncSetAttribute(radix_scatter<KeyT, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  VAPC_RETURN_IF_CUDA(cudaFuncSetAttribute(radix_scatter<KeyT, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  if (passes == 0) {
    if (idx_is_iota) VAPC_LAUNCH(iota_kernel, kNumSMs * 8, 256, 0, s, idx, n);
    *result_location = 0;
    VAPC_CHECK_LAUNCH();
    return VAPC_OK;
  }
  // key buffers: keys -> alt -> (alt2 | keys) -> alt -> ...; with alt2 the caller's keys stay intact
  KeyT* src_k = keys; KeyT* dst_k = keys_alt;
  KeyT* spare = keys_alt2 ? keys_alt2 : keys;
  uint32_t* src_i = idx; uint32_t* dst_i = idx_alt;
  for (int p = 0; p < passes; ++p) {
    const int shift = p * kRadixBits;
    const int bits = min(kRadixBits, end_bit - shift);
    const unsigned mask = (1u << bits) - 1u;
    VAPC_LAUNCH(radix_hist<KeyT>, (unsigned)ntiles, kThreads, 0, s, src_k, n, shift, mask, tile_hist);
    VAPC_LAUNCH(colscan_partial, (unsigned)nchunks, kThreads, 0, s, tile_hist, (int)ntiles, chunk_sum);
    VAPC_LAUNCH(colscan_chunks, 1, kThreads, 0, s, chunk_sum, chunk_base, (int)nchunks);
    VAPC_LAUNCH(colscan_apply, (unsigned)nchunks, kThreads, 0, s, tile_hist, (int)ntiles, chunk_base);
    if (p == 0 && idx_is_iota)
      VAPC_LAUNCH((radix_scatter<KeyT, true>), (unsigned)ntiles, kThreads, smem, s, src_k, nullptr, dst_k, dst_i, n, shift, mask, tile_hist);
    else
      VAPC_LAUNCH((radix_scatter<KeyT, false>), (unsigned)ntiles, kThreads, smem, s, src_k, src_i, dst_k, dst_i, n, shift, mask, tile_hist);
    KeyT* next_dst = (p == 0) ? spare : src_k;
    src_k = dst_k; dst_k = next_dst;
    uint32_t* ti = src_i; src_i = dst_i; dst_i = ti;
  }
  VAPC_CHECK_LAUNCH();
  const int key_loc = src_k == keys ? 0 : (src_k == keys_alt ? 1 : 2);
  *result_location = key_loc | ((passes & 1) ? 4 : 0);
  return VAPC_OK;
}
}  // namespace

extern "C" size_t vapc_sort_workspace_bytes(int64_t n) {
  if (n < 1) n = 1;
  const int64_t ntiles = num_tiles(n);
  return (size_t)(ntiles + 2 * num_chunks(ntiles)) * kRadix * sizeof(uint32_t);
}

extern "C" int vapc_sort_pairs(void* keys, void* keys_alt, void* keys_alt2, uint32_t* idx, uint32_t* idx_alt, int32_t idx_is_iota,
                               int64_t n, int32_t key_bytes, int32_t end_bit, void* ws, size_t ws_bytes, int32_t* result_location_host,
                               void* stream) {
  if (n < 0 || !result_location_host) return vapc_set_error(VAPC_E_INVALID, "vapc_sort_pairs: bad arguments");
  if (n >= (int64_t)1 << 32) return vapc_set_error(VAPC_E_INVALID, "vapc_sort_pairs: n must be < 2^32 per GPU");
  *result_location_host = 0;
  if (n == 0) return VAPC_OK;
  if (!keys || !keys_alt || !idx || !idx_alt) return vapc_set_error(VAPC_E_INVALID, "vapc_sort_pairs: null buffer");
  if (end_bit < 0 || end_bit > key_bytes * 8) return vapc_set_error(VAPC_E_INVALID, "vapc_sort_pairs: end_bit out of range");
  if (ws_bytes < vapc_sort_workspace_bytes(n) || !ws) return vapc_set_error(VAPC_E_WORKSPACE, "vapc_sort_pairs: workspace too small");
  if (key_bytes == 4)
    return sort_impl<uint32_t>(static_cast<uint32_t*>(keys), static_cast<uint32_t*>(keys_alt), static_cast<uint32_t*>(keys_alt2), idx, idx_alt,
                               idx_is_iota, n, end_bit, ws, result_location_host, as_stream(stream));
  if (key_bytes == 8)
    return sort_impl<unsigned long long>(static_cast<unsigned long long*>(keys), static_cast<unsigned long long*>(keys_alt),
                                         static_cast<unsigned long long*>(keys_alt2), idx, idx_alt, idx_is_iota, n, end_bit, ws,
                                         result_location_host, as_stream(stream));
  return vapc_set_error(VAPC_E_INVALID, "vapc_sort_pairs: key_bytes must be 4 or 8");
}
