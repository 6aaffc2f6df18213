// Point -> voxel coordinates -> packed key kernels (streaming, HBM-bound).
//   vapc_minmax_f64        24 B/pt read
//   vapc_voxel_coords_f64  24 B/pt read + 24 B/pt write          (Vapc.voxelize, vapc.py:188-207)
//   vapc_pack_keys_f64     24 B/pt read + key (4|8) + 32 B/pt xyzw record write
#include <math.h>
#include <stdarg.h>
#include <algorithm>
#include <string.h>

#include "common.cuh"

// ------------------------------------------------------------------------------------------
// error plumbing
// ------------------------------------------------------------------------------------------
static thread_local char g_err[512] = "";

int vapc_set_error(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return code;
}
int vapc_cuda_error(cudaError_t e, const char* file, int line) {
  snprintf(g_err, sizeof(g_err), "CUDA error %d (%s) at %s:%d", (int)e, cudaGetErrorString(e), file, line);
  return VAPC_E_CUDA;
}
extern "C" const char* vapc_last_error(void) { return g_err; }

#include <atomic>
static std::atomic<unsigned long long> g_launches{0};
void vapc_note_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }
extern "C" uint64_t vapc_launch_count(void) { return g_launches.load(std::memory_order_relaxed); }
extern "C" int vapc_abi_version(void) { return VAPC_ABI_VERSION; }

// ---- optional per-kernel timing (CUDA events on the launching stream) ------------------------------------
#include <map>
#include <mutex>
#include <string>
#include <vector>
bool g_vapc_profile = false;
namespace {
struct ProfRec { const char* name; cudaEvent_t a, b; };
std::mutex g_prof_mu;
std::vector<ProfRec> g_prof;
}  // namespace
void vapc_profile_begin(const char* name, cudaStream_t s) {
  ProfRec r{name, nullptr, nullptr};
  if (cudaEventCreate(&r.a) != cudaSuccess || cudaEventCreate(&r.b) != cudaSuccess) return;
  cudaEventRecord(r.a, s);
  std::lock_guard<std::mutex> lk(g_prof_mu);
  g_prof.push_back(r);
}
void vapc_profile_end(cudaStream_t s) {
  std::lock_guard<std::mutex> lk(g_prof_mu);
  if (!g_prof.empty()) cudaEventRecord(g_prof.back().b, s);
}
extern "C" int vapc_profile_enable(int32_t on) {
  g_vapc_profile = on != 0;
  return VAPC_OK;
}
// "name<TAB>launches<TAB>total_ms\n" per kernel since the last call; synchronises on the recorded events.
extern "C" int64_t vapc_profile_collect(char* out, int64_t cap) {
  std::lock_guard<std::mutex> lk(g_prof_mu);
  std::map<std::string, std::pair<long, double>> agg;
  for (auto& r : g_prof) {
    float ms = 0.f;
    if (r.a && r.b && cudaEventSynchronize(r.b) == cudaSuccess && cudaEventElapsedTime(&ms, r.a, r.b) == cudaSuccess) {
      auto& e = agg[r.name];
      e.first += 1;
      e.second += ms;
    }
    if (r.a) cudaEventDestroy(r.a);
    if (r.b) cudaEventDestroy(r.b);
  }
  g_prof.clear();
  std::string text;
  for (auto& kv : agg) text += kv.first + "\t" + std::to_string(kv.second.first) + "\t" + std::to_string(kv.second.second) + "\n";
  if (out && cap > 0) {
    const int64_t m = std::min<int64_t>((int64_t)text.size(), cap - 1);
    memcpy(out, text.data(), (size_t)m);
    out[m] = 0;
  }
  return (int64_t)text.size();
}

// The per-voxel gathers read one 32-byte record per point at random; ask the L2 to fetch no more than that
// from DRAM on a miss (the default fetch granularity pulls neighbouring sectors that are never used).
extern "C" int vapc_set_l2_fetch_granularity(int32_t bytes) {
  if (bytes != 32 && bytes != 64 && bytes != 128) return vapc_set_error(VAPC_E_INVALID, "vapc_set_l2_fetch_granularity: 32, 64 or 128");
  VAPC_RETURN_IF_CUDA(cudaDeviceSetLimit(cudaLimitMaxL2FetchGranularity, (size_t)bytes));
  return VAPC_OK;
}

namespace {
constexpr int kThreads = 256;

inline int stream_grid(int64_t n, int per_block) {
  int64_t blocks = (n + per_block - 1) / per_block;
  const int64_t cap = (int64_t)kNumSMs * 16;  // persistent grid-stride: 16 CTAs of 256 threads per SM
  if (blocks > cap) blocks = cap;
  return (int)(blocks < 1 ? 1 : blocks);
}

// ---- min / max -----------------------------------------------------------------------------
template <typename Src>
__global__ void __launch_bounds__(kThreads) minmax_partial(Src src, int64_t n, double* __restrict__ partial /*[grid][6]*/) {
  double mn[3] = {INFINITY, INFINITY, INFINITY}, mx[3] = {-INFINITY, -INFINITY, -INFINITY};
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    const double a = src.cx(i), b = src.cy(i), c = src.cz(i);
    mn[0] = fmin(mn[0], a); mx[0] = fmax(mx[0], a);
    mn[1] = fmin(mn[1], b); mx[1] = fmax(mx[1], b);
    mn[2] = fmin(mn[2], c); mx[2] = fmax(mx[2], c);
  }
  __shared__ double s[6][kThreads / 32];
#pragma unroll
  for (int k = 0; k < 3; ++k) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      mn[k] = fmin(mn[k], __shfl_xor_sync(0xffffffffu, mn[k], o));
      mx[k] = fmax(mx[k], __shfl_xor_sync(0xffffffffu, mx[k], o));
    }
  }
  if ((threadIdx.x & 31) == 0) {
#pragma unroll
    for (int k = 0; k < 3; ++k) { s[k][threadIdx.x >> 5] = mn[k]; s[3 + k][threadIdx.x >> 5] = mx[k]; }
  }
  __syncthreads();
  if (threadIdx.x < 6) {
    double r = s[threadIdx.x][0];
    for (int w = 1; w < kThreads / 32; ++w) r = threadIdx.x < 3 ? fmin(r, s[threadIdx.x][w]) : fmax(r, s[threadIdx.x][w]);
    partial[(int64_t)blockIdx.x * 6 + threadIdx.x] = r;
  }
}
__global__ void minmax_final(const double* __restrict__ partial, int nparts, double* __restrict__ out6) {
  const int k = threadIdx.x >> 5, lane = threadIdx.x & 31;  // 6 warps, one per output
  if (k >= 6) return;
  double r = k < 3 ? INFINITY : -INFINITY;
  for (int i = lane; i < nparts; i += 32) r = k < 3 ? fmin(r, partial[i * 6 + k]) : fmax(r, partial[i * 6 + k]);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const double t = __shfl_xor_sync(0xffffffffu, r, o);
    r = k < 3 ? fmin(r, t) : fmax(r, t);
  }
  if (lane == 0) out6[k] = r;
}

// ---- voxel coordinates ----------------------------------------------------------------------
__global__ void __launch_bounds__(kThreads) voxel_coords_kernel(const double* __restrict__ x, const double* __restrict__ y,
                                                                const double* __restrict__ z, int64_t n, Grid g,
                                                                long long* __restrict__ vx, long long* __restrict__ vy,
                                                                long long* __restrict__ vz) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    __stcs(vx + i, voxel_of_fast(ld_stream(x + i), g.ox, g.vs, g.ivs));
    __stcs(vy + i, voxel_of_fast(ld_stream(y + i), g.oy, g.vs, g.ivs));
    __stcs(vz + i, voxel_of_fast(ld_stream(z + i), g.oz, g.vs, g.ivs));
  }
}

// ---- packed keys (+ 32-byte xyzw records) ---------------------------------------------------
// Each thread handles two consecutive points per iteration so that x/y/z loads are 16-byte
// vector loads when the columns are 16-byte aligned (VEC=2); VEC=1 is the unaligned fallback.
template <typename KeyT, int VEC>
__global__ void __launch_bounds__(kThreads) pack_keys_kernel(const double* __restrict__ x, const double* __restrict__ y,
                                                             const double* __restrict__ z, int64_t n, Grid g,
                                                             KeyT* __restrict__ keys, double4* __restrict__ xyzw,
                                                             int* __restrict__ status) {
  int bad = 0;
  const int64_t nvec = n / VEC;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t v = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; v < nvec; v += stride) {
    if (VEC == 2) {
      const double2 a = __ldcs(reinterpret_cast<const double2*>(x) + v);
      const double2 b = __ldcs(reinterpret_cast<const double2*>(y) + v);
      const double2 c = __ldcs(reinterpret_cast<const double2*>(z) + v);
      const KeyT k0 = pack_one<KeyT>(a.x, b.x, c.x, g, bad);
      const KeyT k1 = pack_one<KeyT>(a.y, b.y, c.y, g, bad);
      if (sizeof(KeyT) == 4) {
        reinterpret_cast<uint2*>(keys)[v] = make_uint2((unsigned)k0, (unsigned)k1);
      } else {
        reinterpret_cast<ulonglong2*>(keys)[v] = make_ulonglong2(k0, k1);
      }
      if (xyzw) {
        st_record(xyzw + 2 * v, a.x, b.x, c.x, index_as_record_w((uint32_t)(2 * v)));
        st_record(xyzw + 2 * v + 1, a.y, b.y, c.y, index_as_record_w((uint32_t)(2 * v + 1)));
      }
    } else {
      const double a = ld_stream(x + v), b = ld_stream(y + v), c = ld_stream(z + v);
      keys[v] = pack_one<KeyT>(a, b, c, g, bad);
      if (xyzw) st_record(xyzw + v, a, b, c, index_as_record_w((uint32_t)v));
    }
  }
  if (VEC == 2 && (n & 1) && blockIdx.x == 0 && threadIdx.x == 0) {
    const int64_t i = n - 1;
    const double a = x[i], b = y[i], c = z[i];
    keys[i] = pack_one<KeyT>(a, b, c, g, bad);
    if (xyzw) xyzw[i] = make_double4(a, b, c, index_as_record_w((uint32_t)i));
  }
  if (__any_sync(0xffffffffu, bad) && (threadIdx.x & 31) == 0) atomicOr(status, 1);
}

template <typename KeyT>
__global__ void __launch_bounds__(kThreads) unpack_keys_kernel(const KeyT* __restrict__ keys, int64_t n, Grid g,
                                                               long long* __restrict__ vx, long long* __restrict__ vy,
                                                               long long* __restrict__ vz) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    long long a, b, c;
    unpack_key((unsigned long long)keys[i], g, a, b, c);
    if (vx) vx[i] = a;
    if (vy) vy[i] = b;
    if (vz) vz[i] = c;
  }
}

// centre = ((v * vs) + vs / 2) + origin ; corner = (v * vs) + origin   (vapc.py:902-906, 922-924)
// int64 -> double conversion, then the reference's operation order without contraction.
template <typename KeyT>
__global__ void __launch_bounds__(kThreads) center_corner_kernel(const KeyT* __restrict__ keys, int64_t n, Grid g,
                                                                 double* __restrict__ center, double* __restrict__ corner) {
  const double half = __ddiv_rn(g.vs, 2.0);
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    long long v[3];
    unpack_key((unsigned long long)keys[i], g, v[0], v[1], v[2]);
    const double o[3] = {g.ox, g.oy, g.oz};
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      const double m = __dmul_rn((double)v[k], g.vs);
      if (center) center[(int64_t)k * n + i] = __dadd_rn(__dadd_rn(m, half), o[k]);
      if (corner) corner[(int64_t)k * n + i] = __dadd_rn(m, o[k]);
    }
  }
}

inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }
inline bool aligned32(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 31u) == 0; }
}  // namespace

// ------------------------------------------------------------------------------------------
// C ABI
// ------------------------------------------------------------------------------------------
extern "C" size_t vapc_minmax_workspace_bytes(void) { return (size_t)kNumSMs * 16 * 6 * sizeof(double); }

extern "C" int vapc_minmax_f64(const double* x, const double* y, const double* z, int64_t n, double* out6, void* ws,
                               size_t ws_bytes, void* stream) {
  if (n <= 0 || !x || !y || !z || !out6) return vapc_set_error(VAPC_E_INVALID, "vapc_minmax_f64: empty input or null pointer");
  if (ws_bytes < vapc_minmax_workspace_bytes() || !ws) return vapc_set_error(VAPC_E_WORKSPACE, "vapc_minmax_f64: workspace too small");
  const int grid = stream_grid(n, kThreads * 8);
  VAPC_LAUNCH(minmax_partial<SrcF64>, grid, kThreads, 0, as_stream(stream), SrcF64{x, y, z}, n, static_cast<double*>(ws));
  VAPC_LAUNCH(minmax_final, 1, 192, 0, as_stream(stream), static_cast<double*>(ws), grid, out6);
  VAPC_CHECK_LAUNCH();
  return VAPC_OK;
}

namespace {
__global__ void __launch_bounds__(kThreads) las_coords_kernel(SrcI32 src, int64_t n, double* __restrict__ x, double* __restrict__ y,
                                                              double* __restrict__ z) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    __stcs(x + i, src.cx(i));
    __stcs(y + i, src.cy(i));
    __stcs(z + i, src.cz(i));
  }
}
}  // namespace

extern "C" int vapc_minmax_i32(const int32_t* X, const int32_t* Y, const int32_t* Z, int64_t n, const vapc_las_transform_t* t_host, double* out6,
                               void* ws, size_t ws_bytes, void* stream) {
  if (n <= 0 || !X || !Y || !Z || !out6 || !t_host) return vapc_set_error(VAPC_E_INVALID, "vapc_minmax_i32: empty input or null pointer");
  if (ws_bytes < vapc_minmax_workspace_bytes() || !ws) return vapc_set_error(VAPC_E_WORKSPACE, "vapc_minmax_i32: workspace too small");
  const int grid = stream_grid(n, kThreads * 8);
  VAPC_LAUNCH(minmax_partial<SrcI32>, grid, kThreads, 0, as_stream(stream), make_src_i32(X, Y, Z, t_host), n, static_cast<double*>(ws));
  VAPC_LAUNCH(minmax_final, 1, 192, 0, as_stream(stream), static_cast<double*>(ws), grid, out6);
  VAPC_CHECK_LAUNCH();
  return VAPC_OK;
}

extern "C" int vapc_las_coords_i32(const int32_t* X, const int32_t* Y, const int32_t* Z, int64_t n, const vapc_las_transform_t* t_host, double* x,
                                   double* y, double* z, void* stream) {
  if (n < 0 || !t_host) return vapc_set_error(VAPC_E_INVALID, "vapc_las_coords_i32: bad arguments");
  if (n == 0) return VAPC_OK;
  if (!X || !Y || !Z || !x || !y || !z) return vapc_set_error(VAPC_E_INVALID, "vapc_las_coords_i32: null buffer");
  VAPC_LAUNCH(las_coords_kernel, stream_grid(n, kThreads * 4), kThreads, 0, as_stream(stream), make_src_i32(X, Y, Z, t_host), n, x, y, z);
  VAPC_CHECK_LAUNCH();
  return VAPC_OK;
}

static int check_grid(const vapc_grid_t* g, const char* who) {
  if (!g) return vapc_set_error(VAPC_E_INVALID, "%s: null grid", who);
  if (!(g->voxel_size > 0)) return vapc_set_error(VAPC_E_INVALID, "%s: voxel_size must be a positive number", who);
  return VAPC_OK;
}

extern "C" int vapc_voxel_coords_f64(const double* x, const double* y, const double* z, int64_t n,
                                     const vapc_grid_t* grid_host, int64_t* vx, int64_t* vy, int64_t* vz, void* stream) {
  if (int e = check_grid(grid_host, "vapc_voxel_coords_f64")) return e;
  if (n < 0) return vapc_set_error(VAPC_E_INVALID, "vapc_voxel_coords_f64: n < 0");
  if (n == 0) return VAPC_OK;
  VAPC_LAUNCH(voxel_coords_kernel, stream_grid(n, kThreads * 4), kThreads, 0, as_stream(stream), 
      x, y, z, n, to_device_grid(grid_host), reinterpret_cast<long long*>(vx), reinterpret_cast<long long*>(vy),
      reinterpret_cast<long long*>(vz));
  VAPC_CHECK_LAUNCH();
  return VAPC_OK;
}

static int bits_for(unsigned long long extent_minus_1) {
  int b = 0;
  while (b < 64 && (extent_minus_1 >> b)) ++b;
  return b;
}

extern "C" int vapc_grid_from_bounds(vapc_grid_t* g, const double* mm) {
  if (int e = check_grid(g, "vapc_grid_from_bounds")) return e;
  if (!mm) return vapc_set_error(VAPC_E_INVALID, "vapc_grid_from_bounds: null bounds");
  int total = 0;
  for (int k = 0; k < 3; ++k) {
    if (!(mm[k] <= mm[3 + k]) || !isfinite(mm[k]) || !isfinite(mm[3 + k]))
      return vapc_set_error(VAPC_E_INVALID, "vapc_grid_from_bounds: non-finite or empty coordinate bounds on axis %d", k);
    // same IEEE expression as the device (host doubles are IEEE; volatile blocks x87/FMA games)
    volatile double lo = (mm[k] - g->origin[k]);
    volatile double hi = (mm[3 + k] - g->origin[k]);
    lo = lo / g->voxel_size;
    hi = hi / g->voxel_size;
    const double flo = floor(lo), fhi = floor(hi);
    if (fabs(flo) > 9.0e18 || fabs(fhi) > 9.0e18)
      return vapc_set_error(VAPC_E_KEYSPACE, "vapc_grid_from_bounds: voxel index overflows int64 on axis %d", k);
    const long long vlo = (long long)flo, vhi = (long long)fhi;
    g->vmin[k] = vlo;
    g->bits[k] = bits_for((unsigned long long)vhi - (unsigned long long)vlo);
    total += g->bits[k];
  }
  if (total > 64)
    return vapc_set_error(VAPC_E_KEYSPACE, "voxel bounding box needs %d key bits (> 64): extent too large for voxel_size", total);
  g->key_bytes = total <= 32 ? 4 : 8;
  return VAPC_OK;
}

extern "C" int vapc_pack_keys_f64(const double* x, const double* y, const double* z, int64_t n,
                                  const vapc_grid_t* grid_host, void* keys, double* xyzw, int32_t* status, void* stream) {
  if (int e = check_grid(grid_host, "vapc_pack_keys_f64")) return e;
  if (n < 0 || !keys || !status) return vapc_set_error(VAPC_E_INVALID, "vapc_pack_keys_f64: bad arguments");
  if (n == 0) return VAPC_OK;
  if (xyzw && !aligned32(xyzw)) return vapc_set_error(VAPC_E_INVALID, "vapc_pack_keys_f64: xyzw must be 32-byte aligned");
  const Grid g = to_device_grid(grid_host);
  const bool vec = aligned16(x) && aligned16(y) && aligned16(z) && aligned16(keys);
  const int grid = stream_grid(n, kThreads * 4);
  cudaStream_t s = as_stream(stream);
  double4* rec = reinterpret_cast<double4*>(xyzw);
  if (grid_host->key_bytes == 4) {
    if (vec) VAPC_LAUNCH((pack_keys_kernel<uint32_t, 2>), grid, kThreads, 0, s, x, y, z, n, g, static_cast<uint32_t*>(keys), rec, status);
    else VAPC_LAUNCH((pack_keys_kernel<uint32_t, 1>), grid, kThreads, 0, s, x, y, z, n, g, static_cast<uint32_t*>(keys), rec, status);
  } else if (grid_host->key_bytes == 8) {
    typedef unsigned long long u64;
    if (vec) VAPC_LAUNCH((pack_keys_kernel<u64, 2>), grid, kThreads, 0, s, x, y, z, n, g, static_cast<u64*>(keys), rec, status);
    else VAPC_LAUNCH((pack_keys_kernel<u64, 1>), grid, kThreads, 0, s, x, y, z, n, g, static_cast<u64*>(keys), rec, status);
  } else {
    return vapc_set_error(VAPC_E_INVALID, "vapc_pack_keys_f64: key_bytes must be 4 or 8");
  }
  VAPC_CHECK_LAUNCH();
  return VAPC_OK;
}

extern "C" int vapc_unpack_keys(const void* keys, int64_t n, const vapc_grid_t* grid_host, int64_t* vx, int64_t* vy,
                                int64_t* vz, void* stream) {
  if (int e = check_grid(grid_host, "vapc_unpack_keys")) return e;
  if (n <= 0) return n == 0 ? VAPC_OK : vapc_set_error(VAPC_E_INVALID, "vapc_unpack_keys: n < 0");
  const Grid g = to_device_grid(grid_host);
  const int grid = stream_grid(n, kThreads * 4);
  if (grid_host->key_bytes == 4)
    VAPC_LAUNCH(unpack_keys_kernel<uint32_t>, grid, kThreads, 0, as_stream(stream), static_cast<const uint32_t*>(keys), n, g,
        reinterpret_cast<long long*>(vx), reinterpret_cast<long long*>(vy), reinterpret_cast<long long*>(vz));
  else
    VAPC_LAUNCH(unpack_keys_kernel<unsigned long long>, grid, kThreads, 0, as_stream(stream), 
        static_cast<const unsigned long long*>(keys), n, g, reinterpret_cast<long long*>(vx),
        reinterpret_cast<long long*>(vy), reinterpret_cast<long long*>(vz));
  VAPC_CHECK_LAUNCH();
  return VAPC_OK;
}

extern "C" int vapc_voxel_center_corner(const void* voxel_keys, int64_t nvoxels, const vapc_grid_t* grid_host,
                                        double* center3, double* corner3, void* stream) {
  if (int e = check_grid(grid_host, "vapc_voxel_center_corner")) return e;
  if (nvoxels <= 0) return nvoxels == 0 ? VAPC_OK : vapc_set_error(VAPC_E_INVALID, "vapc_voxel_center_corner: n < 0");
  const Grid g = to_device_grid(grid_host);
  const int grid = stream_grid(nvoxels, kThreads * 4);
  if (grid_host->key_bytes == 4)
    VAPC_LAUNCH(center_corner_kernel<uint32_t>, grid, kThreads, 0, as_stream(stream), static_cast<const uint32_t*>(voxel_keys), nvoxels, g, center3, corner3);
  else
    VAPC_LAUNCH(center_corner_kernel<unsigned long long>, grid, kThreads, 0, as_stream(stream), static_cast<const unsigned long long*>(voxel_keys), nvoxels, g, center3, corner3);
  VAPC_CHECK_LAUNCH();
  return VAPC_OK;
}
