// Shared device/host helpers for libvapc_b200 (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/vapc_b200.h"

// ---- error plumbing (thread-local message behind vapc_last_error) -----------------------
int vapc_set_error(int code, const char* fmt, ...);
int vapc_cuda_error(cudaError_t e, const char* file, int line);

#define VAPC_RETURN_IF_CUDA(expr)                                        \
  do {                                                                   \
    cudaError_t _e = (expr);                                             \
    if (_e != cudaSuccess) return vapc_cuda_error(_e, __FILE__, __LINE__); \
  } while (0)
#define VAPC_CHECK_LAUNCH() VAPC_RETURN_IF_CUDA(cudaGetLastError())

// every kernel launch of the library goes through this so that vapc_launch_count() is exact; with
// vapc_profile_enable(1) every launch is also bracketed by CUDA events on its own stream (bench.py's per-kernel roofline)
void vapc_note_launch();
extern bool g_vapc_profile;
void vapc_profile_begin(const char* name, cudaStream_t s);
void vapc_profile_end(cudaStream_t s);
#define VAPC_LAUNCH(kernel, grid, block, smem, stream, ...)      \
  do {                                                           \
    if (g_vapc_profile) vapc_profile_begin(#kernel, (stream));   \
    kernel<<<(grid), (block), (smem), (stream)>>>(__VA_ARGS__);  \
    if (g_vapc_profile) vapc_profile_end((stream));              \
    vapc_note_launch();                                          \
  } while (0)

static inline cudaStream_t as_stream(void* s) { return reinterpret_cast<cudaStream_t>(s); }

constexpr int kNumSMs = 148;  // B200: 2 dies x 74 SMs; grids are sized in multiples of this

// ---- packed-key layout, passed by value to kernels -------------------------------------
struct Grid {
  double ox, oy, oz, vs;
  double ivs;  // 1 / vs, correctly rounded (host): the fast path of voxel_of_fast
  double vs3;  // vs ** 3 as the host's pow forms it — what Python's `voxel_size ** 3` is (vapc.py:740); vs * vs * vs is one ulp off for
               // a quarter of all voxel sizes (0.3 among them), and point_density would differ from the reference's in its last bit
  long long minx, miny, minz;
  int bx, by, bz;
};
static inline Grid to_device_grid(const vapc_grid_t* g) {
  Grid d;
  d.ox = g->origin[0]; d.oy = g->origin[1]; d.oz = g->origin[2];
  d.vs = g->voxel_size;
  d.ivs = 1.0 / g->voxel_size;
  d.vs3 = pow(g->voxel_size, 3.0);
  d.minx = g->vmin[0]; d.miny = g->vmin[1]; d.minz = g->vmin[2];
  d.bx = g->bits[0]; d.by = g->bits[1]; d.bz = g->bits[2];
  return d;
}

// v = floor((c - o) / vs): IEEE subtract, IEEE divide (no reciprocal, no FMA contraction),
// floor, convert.  Bit-exact with numpy's np.floor((X - o) / vs).astype(int)  (vapc.py:199-201).
__device__ __forceinline__ long long voxel_of(double c, double o, double vs) {
  return __double2ll_rd(__ddiv_rn(__dsub_rn(c, o), vs));
}

// The same integer without the division for almost every point (an IEEE fp64 division is ~30 instructions, and the key kernels
// form three per point).  With d = c - o (IEEE), Q = d / vs (real) and ivs = fl(1 / vs):
//   |fl(d / vs) - Q| <= 2^-53 |Q|   and   q' = fl(d * ivs) = Q (1 + e1)(1 + e2), |e| <= 2^-53   =>   |q' - fl(d / vs)| < 2^-51 |q'|.
// So when no integer lies within 2^-50 |q'| of q', fl(d / vs) lies strictly between the same two integers as q' and
// floor(q') == floor(fl(d / vs)).  Otherwise (a coordinate on a voxel face within rounding — LAS-quantised clouds have ~0.1 % of
// those —, q' == 0, |q'| >= 2^52) the division itself decides.  NaN / inf fall through to the conversion, which treats them like
// the division's result.  Bit-exact with voxel_of by construction; tests/test_gpu_kernels.py::test_voxel_coords_bit_exact checks
// 80 000 face-adjacent values per grid through the key kernels.
__device__ __forceinline__ long long voxel_of_fast(double c, double o, double vs, double ivs) {
  const double d = __dsub_rn(c, o);
  const double q = __dmul_rn(d, ivs);
  if (fabs(__dsub_rn(q, rint(q))) <= __dmul_rn(fabs(q), 8.8817841970012523e-16 /* 2^-50 */)) return __double2ll_rd(__ddiv_rn(d, vs));
  return __double2ll_rd(q);
}

__device__ __forceinline__ void unpack_key(unsigned long long key, const Grid& g, long long& vx,
                                           long long& vy, long long& vz) {
  const unsigned long long mz = g.bz >= 64 ? ~0ull : ((1ull << g.bz) - 1);
  const unsigned long long my = g.by >= 64 ? ~0ull : ((1ull << g.by) - 1);
  vz = (long long)(key & mz) + g.minz;
  vy = (long long)((g.bz >= 64 ? 0ull : (key >> g.bz)) & my) + g.miny;
  const int s = g.by + g.bz;
  vx = (long long)(s >= 64 ? 0ull : (key >> s)) + g.minx;
}

// Shift point c(v) = origin + (v + 0.5) * vs used for the shifted moments.  It only has to be
// the SAME value wherever a voxel's moments are formed or unshifted, so the operation order
// is pinned with explicit round-to-nearest intrinsics (no contraction).
__device__ __forceinline__ double shift_center(long long v, double o, double vs) {
  return __dadd_rn(o, __dmul_rn(__dadd_rn((double)v, 0.5), vs));
}

// Packed key of one point (x most significant).  `bad` is OR-ed with 1 when a relative voxel coordinate does
// not fit its bit field (NaN / out-of-box point).
template <typename KeyT>
__device__ __forceinline__ KeyT pack_one(double px, double py, double pz, const Grid& g, int& bad, unsigned long long* rx_out = nullptr) {
  const unsigned long long rx = (unsigned long long)(voxel_of_fast(px, g.ox, g.vs, g.ivs) - g.minx);
  const unsigned long long ry = (unsigned long long)(voxel_of_fast(py, g.oy, g.vs, g.ivs) - g.miny);
  const unsigned long long rz = (unsigned long long)(voxel_of_fast(pz, g.oz, g.vs, g.ivs) - g.minz);
  // (unsigned) relative coordinates must fit their bit fields; NaN / out-of-box points do not
  bad |= (g.bx < 64 && (rx >> g.bx)) || (g.by < 64 && (ry >> g.by)) || (g.bz < 64 && (rz >> g.bz));
  const int sxy = g.by + g.bz;
  unsigned long long k = rz;
  if (g.bz < 64) k |= ry << g.bz;
  if (sxy < 64) k |= rx << sxy;
  if (rx_out) *rx_out = rx;
  return (KeyT)k;
}

template <typename T>
__device__ __forceinline__ T ld_stream(const T* p) { return __ldcs(p); }

// One 32-byte xyzw record with a single 256-bit load (LDG.E.256 on sm_100a): one request, one sector.
__device__ __forceinline__ double4 ld_record(const double4* p) {
  double4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.f64 {%0, %1, %2, %3}, [%4];" : "=d"(r.x), "=d"(r.y), "=d"(r.z), "=d"(r.w) : "l"(p));
  return r;
}
// the original point index travels in the record's fourth slot as a bit pattern
__device__ __forceinline__ double index_as_record_w(uint32_t i) { return __longlong_as_double((long long)i); }
__device__ __forceinline__ uint32_t record_index(const double4& r) { return (uint32_t)__double_as_longlong(r.w); }
__device__ __forceinline__ void st_record(double4* p, double x, double y, double z, double w) {
  asm volatile("st.global.L1::no_allocate.v4.f64 [%0], {%1, %2, %3, %4};" ::"l"(p), "d"(x), "d"(y), "d"(z), "d"(w) : "memory");
}

// ---- coordinate sources of the key kernels ------------------------------------------------------------------
// SrcF64: the fp64 columns of the frame.  SrcI32: the integer coordinates of a LAS file with the header's scale / offset,
// coordinate = X * scale + offset evaluated like laspy / numpy evaluate it (one IEEE multiply, one IEEE add, no FMA;
// datahandler.py:72-95 receives exactly these doubles): 12 B/point over PCIe and from HBM instead of 24.
struct SrcF64 {
  const double *x, *y, *z;
  static constexpr unsigned kPairAlign = 15u;  // a pair of points is one 16-byte load
  __device__ __forceinline__ double cx(int64_t i) const { return ld_stream(x + i); }
  __device__ __forceinline__ double cy(int64_t i) const { return ld_stream(y + i); }
  __device__ __forceinline__ double cz(int64_t i) const { return ld_stream(z + i); }
  __device__ __forceinline__ void pair(int64_t v, double (&a)[2], double (&b)[2], double (&c)[2]) const {
    const double2 p = __ldcs(reinterpret_cast<const double2*>(x) + v), q = __ldcs(reinterpret_cast<const double2*>(y) + v),
                  r = __ldcs(reinterpret_cast<const double2*>(z) + v);
    a[0] = p.x; a[1] = p.y; b[0] = q.x; b[1] = q.y; c[0] = r.x; c[1] = r.y;
  }
  bool aligned() const { return ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(y) | reinterpret_cast<uintptr_t>(z)) & kPairAlign) == 0; }
  bool null() const { return !x || !y || !z; }
};
struct SrcI32 {
  const int32_t *x, *y, *z;
  double sx, sy, sz, ox, oy, oz;
  static constexpr unsigned kPairAlign = 7u;
  static __device__ __forceinline__ double conv(int32_t q, double s, double o) { return __dadd_rn(__dmul_rn((double)q, s), o); }
  __device__ __forceinline__ double cx(int64_t i) const { return conv(__ldcs(x + i), sx, ox); }
  __device__ __forceinline__ double cy(int64_t i) const { return conv(__ldcs(y + i), sy, oy); }
  __device__ __forceinline__ double cz(int64_t i) const { return conv(__ldcs(z + i), sz, oz); }
  __device__ __forceinline__ void pair(int64_t v, double (&a)[2], double (&b)[2], double (&c)[2]) const {
    const int2 p = __ldcs(reinterpret_cast<const int2*>(x) + v), q = __ldcs(reinterpret_cast<const int2*>(y) + v),
               r = __ldcs(reinterpret_cast<const int2*>(z) + v);
    a[0] = conv(p.x, sx, ox); a[1] = conv(p.y, sx, ox); b[0] = conv(q.x, sy, oy); b[1] = conv(q.y, sy, oy);
    c[0] = conv(r.x, sz, oz); c[1] = conv(r.y, sz, oz);
  }
  bool aligned() const { return ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(y) | reinterpret_cast<uintptr_t>(z)) & kPairAlign) == 0; }
  bool null() const { return !x || !y || !z; }
};
static inline SrcI32 make_src_i32(const int32_t* x, const int32_t* y, const int32_t* z, const vapc_las_transform_t* t) {
  return SrcI32{x, y, z, t->scale[0], t->scale[1], t->scale[2], t->offset[0], t->offset[1], t->offset[2]};
}

// ---- block-wide exclusive scan of one unsigned per thread (blockDim.x multiple of 32, <= 1024)
// smem33: 33 words of shared scratch.  Returns the exclusive prefix; *total = block sum (all
// threads).  Ends with a barrier, so smem33 may be reused immediately.
__device__ __forceinline__ unsigned block_exclusive_scan(unsigned v, unsigned* smem33, unsigned* total) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  unsigned inc = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const unsigned t = __shfl_up_sync(0xffffffffu, inc, o);
    if (lane >= o) inc += t;
  }
  if (lane == 31) smem33[warp] = inc;
  __syncthreads();
  if (warp == 0) {
    const int nw = (blockDim.x + 31) >> 5;
    const unsigned w = lane < nw ? smem33[lane] : 0u;
    unsigned winc = w;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const unsigned t = __shfl_up_sync(0xffffffffu, winc, o);
      if (lane >= o) winc += t;
    }
    smem33[lane] = winc - w;  // exclusive prefix of the warp sums
    if (lane == 31) smem33[32] = winc;
  }
  __syncthreads();
  const unsigned res = inc - v + smem33[warp];
  *total = smem33[32];
  __syncthreads();
  return res;
}

// ---- device-wide exclusive scan of uint32 (three launches; scan.cu) ---------------------
size_t scan_u32_workspace_bytes(int64_t n);
// out may alias in.  total (nullable, device uint64) receives the grand total.
cudaError_t scan_u32_exclusive(const uint32_t* in, uint32_t* out, int64_t n, unsigned long long* total,
                               void* ws, cudaStream_t s);
