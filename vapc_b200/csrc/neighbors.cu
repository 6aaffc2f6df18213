// Neighbour queries on the voxel table used as a cell list: SURVEY 8(f) row 4.
//   * vapc_cell_nn            exact nearest neighbour of every query point in a voxelised point set
//                             (bi_vapc.py:51-74, 173-186: the two KD-tree queries of mutual_nearest_neighbors)
//   * vapc_cluster_components connected components of the "within cluster_distance" graph
//                             (vapc.py:1274-1357: KD-tree region growing with label merging)
// A VoxelCloud built over the searched points with cell size h IS the cell list: voxel_keys (sorted, x most
// significant), start[], pos_sorted[] and the 32-byte records {x, y, z, original index}.  The cells of one (x, y)
// column and a z range are a contiguous run of keys, so a cube of cells is visited with one binary search per column.
// Every per-element routine is __host__ __device__: the kernels call it per thread, the vapc_selftest_*_host entry
// points run the same code serially on host arrays (CPU tests of the traversal logic, no GPU needed).
#include <math.h>

#include "common.cuh"

namespace {
constexpr int kThreads = 128;
#define HD __host__ __device__ __forceinline__

template <typename KeyT>
struct CellList {
  Grid g;
  const KeyT* keys;       // [ncells] sorted cell keys
  long long ncells;
  const uint32_t* start;  // [ncells + 1] first sorted position of each cell
  const uint32_t* pos;    // [n] sorted position -> record
  const double4* rec;     // [n] {x, y, z, original index as bit pattern}
  long long n;
};

HD long long cell_of(double c, double o, double vs) {
#ifdef __CUDA_ARCH__
  return voxel_of(c, o, vs);
#else
  return (long long)floor((c - o) / vs);
#endif
}
HD uint32_t rec_index(const double4& r) {
#ifdef __CUDA_ARCH__
  return (uint32_t)__double_as_longlong(r.w);
#else
  long long v;
  memcpy(&v, &r.w, 8);
  return (uint32_t)v;
#endif
}
// squared distance exactly as scipy's KDTree forms it for 3 columns: plain products summed left to right, no FMA
HD double dist2(double ax, double ay, double az, const double4& b) {
  const double dx = ax - b.x, dy = ay - b.y, dz = az - b.z;
#ifdef __CUDA_ARCH__
  return __dadd_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)), __dmul_rn(dz, dz));
#else
  volatile double xx = dx * dx, yy = dy * dy, zz = dz * dz;  // volatile: keep the host compiler from contracting
  volatile double s = xx + yy;
  return s + zz;
#endif
}
HD double root_of(double v) {
#ifdef __CUDA_ARCH__
  return __dsqrt_rn(v);
#else
  return sqrt(v);
#endif
}

HD long long axis_cells(int bits) { return bits >= 62 ? (1ll << 62) : (1ll << bits); }

template <typename KeyT>
HD KeyT rel_key(const Grid& g, long long rx, long long ry, long long rz) {
  unsigned long long k = (unsigned long long)rz;
  if (g.bz < 64) k |= (unsigned long long)ry << g.bz;
  if (g.by + g.bz < 64) k |= (unsigned long long)rx << (g.by + g.bz);
  return (KeyT)k;
}

template <typename KeyT>
HD long long lower_bound_key(const KeyT* keys, long long n, KeyT k) {
  long long lo = 0, hi = n;
  while (lo < hi) {
    const long long mid = (lo + hi) >> 1;
    if (keys[mid] < k) lo = mid + 1; else hi = mid;
  }
  return lo;
}

// Calls f(record) for every point whose cell lies in the cube [c - r, c + r]^3 (clipped to the grid).  When the cube has
// more columns than a linear pass over all records costs, that pass is made instead — or, with allow_linear == false,
// nothing is visited and kRefused is returned (the caller hands the query to the tiled all-pairs kernel).  kAll: every
// point of the set has been visited (linear pass, or the cube covers the whole grid).
enum { kVisited = 0, kAll = 1, kRefused = 2 };
template <typename KeyT, typename F>
HD int visit_cube(const CellList<KeyT>& cl, long long cx, long long cy, long long cz, long long r, F&& f, bool allow_linear = true) {
  const Grid& g = cl.g;
  const long long ex = axis_cells(g.bx), ey = axis_cells(g.by), ez = axis_cells(g.bz);
  long long x0 = cx - r - g.minx, x1 = cx + r - g.minx;
  long long y0 = cy - r - g.miny, y1 = cy + r - g.miny;
  long long z0 = cz - r - g.minz, z1 = cz + r - g.minz;
  const bool all = x0 <= 0 && y0 <= 0 && z0 <= 0 && x1 >= ex - 1 && y1 >= ey - 1 && z1 >= ez - 1;
  x0 = x0 < 0 ? 0 : x0; y0 = y0 < 0 ? 0 : y0; z0 = z0 < 0 ? 0 : z0;
  x1 = x1 > ex - 1 ? ex - 1 : x1; y1 = y1 > ey - 1 ? ey - 1 : y1; z1 = z1 > ez - 1 ? ez - 1 : z1;
  if (x0 > x1 || y0 > y1 || z0 > z1) return all ? kAll : kVisited;
  const double columns = (double)(x1 - x0 + 1) * (double)(y1 - y0 + 1);
  if (all || columns * 24.0 > (double)cl.n) {
    if (!allow_linear && cl.n > 4096) return kRefused;
    for (long long p = 0; p < cl.n; ++p) f(cl.rec[p]);
    return kAll;
  }
  for (long long x = x0; x <= x1; ++x) {
    for (long long y = y0; y <= y1; ++y) {
      const KeyT klo = rel_key<KeyT>(g, x, y, z0), khi = rel_key<KeyT>(g, x, y, z1);
      long long v = lower_bound_key(cl.keys, cl.ncells, klo);
      if (v >= cl.ncells || cl.keys[v] > khi) continue;
      const uint32_t p0 = cl.start[v];
      while (v < cl.ncells && cl.keys[v] <= khi) ++v;  // the cells of the run are adjacent rows
      const uint32_t p1 = cl.start[v];
      for (uint32_t p = p0; p < p1; ++p) f(cl.rec[cl.pos[p]]);
    }
  }
  return kVisited;
}

// ---- exact nearest neighbour ------------------------------------------------------------------------------
// Cube of radius 1 first; a candidate at distance <= r h is final (everything outside the cube is farther than
// r h); otherwise one more cube of radius ceil(best / h) holds the true neighbour.  Nothing found: double r.
// Ties: the lowest original index (a KD-tree's choice among exactly equidistant points is traversal order).
// bounded: the grid search gives up (index 0xffffffff) when the neighbour is farther than ~8 cells — a grid degenerates
// to a scan of the whole set per query there; those queries go to the tiled all-pairs kernel, whose cost is bounded.
constexpr long long kMaxEmptyRadius = 4, kMaxFinalRadius = 8;
template <typename KeyT>
HD void cell_nn_one(const CellList<KeyT>& cl, double qx, double qy, double qz, bool bounded, uint32_t* idx_out, double* dist_out) {
  const Grid& g = cl.g;
  const long long cx = cell_of(qx, g.ox, g.vs), cy = cell_of(qy, g.oy, g.vs), cz = cell_of(qz, g.oz, g.vs);
  double best2 = INFINITY;
  uint32_t besti = 0xffffffffu;
  auto f = [&](const double4& b) {
    const double d2 = dist2(qx, qy, qz, b);
    const uint32_t j = rec_index(b);
    if (d2 < best2 || (d2 == best2 && j < besti)) { best2 = d2; besti = j; }
  };
  bool refused = false;
  long long r = 1;
  for (int guard = 0; guard < 64; ++guard) {
    const int st = visit_cube(cl, cx, cy, cz, r, f, !bounded);
    if (st == kRefused) { refused = true; break; }
    if (st == kAll) break;
    if (best2 < INFINITY) {
      const double bound = (double)r * g.vs;
      if (best2 <= bound * bound * (1.0 - 1e-9)) break;
      long long big = (long long)ceil(root_of(best2) / g.vs * (1.0 + 1e-9));
      if (big <= r) big = r + 1;
      if (bounded && big > kMaxFinalRadius) { refused = true; break; }
      if (visit_cube(cl, cx, cy, cz, big, f, !bounded) == kRefused) refused = true;
      break;
    }
    if (bounded && r >= kMaxEmptyRadius) { refused = true; break; }
    r = r < (1ll << 40) ? r * 2 : r;
  }
  *idx_out = refused ? 0xffffffffu : besti;
  *dist_out = refused ? INFINITY : root_of(best2);
}

// ---- connected components -----------------------------------------------------------------------------------
HD uint32_t uf_load(const uint32_t* p) {
#ifdef __CUDA_ARCH__
  return *reinterpret_cast<const volatile uint32_t*>(p);
#else
  return *p;
#endif
}
HD uint32_t uf_cas(uint32_t* p, uint32_t expect, uint32_t val) {
#ifdef __CUDA_ARCH__
  return atomicCAS(p, expect, val);
#else
  const uint32_t old = *p;
  if (old == expect) *p = val;
  return old;
#endif
}
// root of x with path halving; parents only ever decrease, so concurrent halving writes are benign
HD uint32_t uf_find(uint32_t* parent, uint32_t x) {
  uint32_t p = uf_load(parent + x);
  while (p != x) {
    const uint32_t gp = uf_load(parent + p);
    if (gp != p) parent[x] = gp;
    x = p;
    p = gp;
  }
  return x;
}
// the larger root is hooked under the smaller one: the root of a finished component is its lowest unit index
HD void uf_union(uint32_t* parent, uint32_t a, uint32_t b) {
  for (;;) {
    a = uf_find(parent, a);
    b = uf_find(parent, b);
    if (a == b) return;
    if (a < b) { const uint32_t t = a; a = b; b = t; }
    const uint32_t old = uf_cas(parent + a, a, b);
    if (old == a) return;
    a = old;
  }
}

// unit at sorted position p: link it with every earlier unit within `dist`, m1 = lowest first[] in its closed
// neighbourhood (vapc.py:1341-1347: the points one KD-tree query labels together)
template <typename KeyT>
HD void cluster_link_one(const CellList<KeyT>& cl, long long p, double dist, const uint32_t* first, uint32_t* parent, uint32_t* m1) {
  const Grid& g = cl.g;
  const double4 a = cl.rec[cl.pos[p]];
  const uint32_t i = rec_index(a);
  uint32_t m = first ? first[i] : i;
  auto f = [&](const double4& b) {
    if (!(root_of(dist2(a.x, a.y, a.z, b)) <= dist)) return;
    const uint32_t j = rec_index(b);
    const uint32_t fj = first ? first[j] : j;
    m = fj < m ? fj : m;
    if (j < i) uf_union(parent, i, j);
  };
  visit_cube(cl, cell_of(a.x, g.ox, g.vs), cell_of(a.y, g.oy, g.vs), cell_of(a.z, g.oz, g.vs), 1, f);
  m1[i] = m;
}
// m2 = lowest first[] within two hops: a unit opens a new label iff nothing within two hops was visited before it
template <typename KeyT>
HD void cluster_hop2_one(const CellList<KeyT>& cl, long long p, double dist, const uint32_t* m1, uint32_t* m2) {
  const Grid& g = cl.g;
  const double4 a = cl.rec[cl.pos[p]];
  const uint32_t i = rec_index(a);
  uint32_t m = m1[i];
  auto f = [&](const double4& b) {
    if (!(root_of(dist2(a.x, a.y, a.z, b)) <= dist)) return;
    const uint32_t mj = m1[rec_index(b)];
    m = mj < m ? mj : m;
  };
  visit_cube(cl, cell_of(a.x, g.ox, g.vs), cell_of(a.y, g.oy, g.vs), cell_of(a.z, g.oz, g.vs), 1, f);
  m2[i] = m;
}

template <typename KeyT>
__global__ void __launch_bounds__(kThreads) cell_nn_kernel(CellList<KeyT> cl, const double* __restrict__ qx, const double* __restrict__ qy,
                                                           const double* __restrict__ qz, long long nq, bool bounded,
                                                           uint32_t* __restrict__ nn_idx, double* __restrict__ nn_dist) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nq) return;
  uint32_t idx;
  double d;
  cell_nn_one(cl, qx[i], qy[i], qz[i], bounded, &idx, &d);
  nn_idx[i] = idx;
  nn_dist[i] = d;
}

// Tiled all-pairs search for the queries the grid search gave up on (which[] lists them): every thread keeps one query,
// the CTA streams the records through shared memory.  Same distance formula, same tie rule.
__global__ void __launch_bounds__(kThreads) nn_brute_kernel(const double4* __restrict__ rec, long long n, const double* __restrict__ qx,
                                                            const double* __restrict__ qy, const double* __restrict__ qz,
                                                            const long long* __restrict__ which, long long nwhich, uint32_t* __restrict__ nn_idx,
                                                            double* __restrict__ nn_dist) {
  __shared__ double4 tile[kThreads];
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const bool active = t < nwhich;
  const long long q = active ? which[t] : 0;
  const double ax = active ? qx[q] : 0.0, ay = active ? qy[q] : 0.0, az = active ? qz[q] : 0.0;
  double best2 = INFINITY;
  uint32_t besti = 0xffffffffu;
  for (long long base = 0; base < n; base += kThreads) {
    const long long p = base + threadIdx.x;
    if (p < n) tile[threadIdx.x] = rec[p];
    __syncthreads();
    const int m = (int)min((long long)kThreads, n - base);
    for (int k = 0; k < m; ++k) {
      const double4 b = tile[k];
      const double d2 = dist2(ax, ay, az, b);
      const uint32_t j = rec_index(b);
      if (d2 < best2 || (d2 == best2 && j < besti)) { best2 = d2; besti = j; }
    }
    __syncthreads();
  }
  if (active) {
    nn_idx[q] = besti;
    nn_dist[q] = root_of(best2);
  }
}

__global__ void __launch_bounds__(256) cluster_init_kernel(uint32_t* parent, uint32_t* comp_first, unsigned long long* comp_size, long long n) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  parent[i] = (uint32_t)i;
  comp_first[i] = 0xffffffffu;
  comp_size[i] = 0ull;
}
template <typename KeyT>
__global__ void __launch_bounds__(kThreads) cluster_link_kernel(CellList<KeyT> cl, double dist, const uint32_t* __restrict__ first,
                                                                uint32_t* parent, uint32_t* __restrict__ m1) {
  const long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (p < cl.n) cluster_link_one(cl, p, dist, first, parent, m1);
}
template <typename KeyT>
__global__ void __launch_bounds__(kThreads) cluster_hop2_kernel(CellList<KeyT> cl, double dist, const uint32_t* __restrict__ m1,
                                                                uint32_t* __restrict__ m2) {
  const long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (p < cl.n) cluster_hop2_one(cl, p, dist, m1, m2);
}
__global__ void __launch_bounds__(256) cluster_finish_kernel(uint32_t* parent, const uint32_t* __restrict__ first, const uint32_t* __restrict__ weight,
                                                             uint32_t* __restrict__ root, uint32_t* comp_first, unsigned long long* comp_size, long long n) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const uint32_t r = uf_find(parent, (uint32_t)i);
  root[i] = r;
  atomicMin(comp_first + r, first ? first[i] : (uint32_t)i);
  atomicAdd(comp_size + r, (unsigned long long)(weight ? weight[i] : 1u));
}


// ---- (dist, index) candidates reduced per row: lexicographic minimum -------------------------------------------
// Distances are >= 0, so their bit patterns order like the values: pass 1 atomicMin of the distance bits, pass 2
// atomicMin of the index among the candidates that attain it.  Integer atomics only: the result does not depend
// on the order of arrival.
__global__ void __launch_bounds__(256) cand_init_kernel(unsigned long long* best_dist, unsigned long long* best_idx, long long nrows) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < nrows) { best_dist[i] = ~0ull; best_idx[i] = ~0ull; }
}
__global__ void __launch_bounds__(256) cand_min_dist_kernel(const long long* __restrict__ rows, const double* __restrict__ dist, long long n,
                                                            unsigned long long* best_dist) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) atomicMin(best_dist + rows[i], (unsigned long long)__double_as_longlong(dist[i]));
}
__global__ void __launch_bounds__(256) cand_min_idx_kernel(const long long* __restrict__ rows, const double* __restrict__ dist,
                                                           const long long* __restrict__ idx, long long n, const unsigned long long* __restrict__ best_dist,
                                                           unsigned long long* best_idx) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n && (unsigned long long)__double_as_longlong(dist[i]) == best_dist[rows[i]]) atomicMin(best_idx + rows[i], (unsigned long long)idx[i]);
}

template <typename KeyT>
CellList<KeyT> make_list(const vapc_grid_t* grid, const void* keys, int64_t ncells, const uint32_t* start, const uint32_t* pos,
                         const double* xyzw, int64_t n) {
  CellList<KeyT> cl;
  cl.g = to_device_grid(grid);
  cl.keys = reinterpret_cast<const KeyT*>(keys);
  cl.ncells = ncells;
  cl.start = start;
  cl.pos = pos;
  cl.rec = reinterpret_cast<const double4*>(xyzw);
  cl.n = n;
  return cl;
}

int check_list(const char* who, const vapc_grid_t* grid, const void* keys, int64_t ncells, const uint32_t* start, const uint32_t* pos,
               const double* xyzw, int64_t n) {
  if (!grid || !keys || !start || !pos || !xyzw) return vapc_set_error(VAPC_E_INVALID, "%s: null cell list", who);
  if (n <= 0 || ncells <= 0 || n >= (1ll << 32)) return vapc_set_error(VAPC_E_INVALID, "%s: the searched set must hold 1 .. 2^32 - 1 points", who);
  if (grid->key_bytes != 4 && grid->key_bytes != 8) return vapc_set_error(VAPC_E_INVALID, "%s: key_bytes must be 4 or 8", who);
  if (!(grid->voxel_size > 0)) return vapc_set_error(VAPC_E_INVALID, "%s: cell size must be positive", who);
  return VAPC_OK;
}

template <typename KeyT>
void cluster_host(const CellList<KeyT>& cl, double dist, const uint32_t* first, const uint32_t* weight, uint32_t* root, uint32_t* hop2,
                  uint32_t* comp_first, int64_t* comp_size, uint32_t* parent, uint32_t* m1) {
  for (long long i = 0; i < cl.n; ++i) { parent[i] = (uint32_t)i; comp_first[i] = 0xffffffffu; comp_size[i] = 0; }
  for (long long p = 0; p < cl.n; ++p) cluster_link_one(cl, p, dist, first, parent, m1);
  for (long long p = 0; p < cl.n; ++p) cluster_hop2_one(cl, p, dist, m1, hop2);
  for (long long i = 0; i < cl.n; ++i) {
    const uint32_t r = uf_find(parent, (uint32_t)i);
    root[i] = r;
    const uint32_t fi = first ? first[i] : (uint32_t)i;
    if (fi < comp_first[r]) comp_first[r] = fi;
    comp_size[r] += weight ? weight[i] : 1u;
  }
}
}  // namespace

extern "C" int vapc_cell_nn(const double* qx, const double* qy, const double* qz, int64_t nq, const vapc_grid_t* grid_host,
                            const void* cell_keys, int64_t ncells, const uint32_t* start, const uint32_t* pos_sorted, const double* xyzw,
                            int64_t n, int32_t bounded, uint32_t* nn_idx, double* nn_dist, void* stream) {
  if (nq < 0) return vapc_set_error(VAPC_E_INVALID, "vapc_cell_nn: negative size");
  if (nq == 0) return VAPC_OK;
  if (const int rc = check_list("vapc_cell_nn", grid_host, cell_keys, ncells, start, pos_sorted, xyzw, n)) return rc;
  if (!qx || !qy || !qz || !nn_idx || !nn_dist) return vapc_set_error(VAPC_E_INVALID, "vapc_cell_nn: null argument");
  const unsigned blocks = (unsigned)((nq + kThreads - 1) / kThreads);
  if (grid_host->key_bytes == 4) {
    VAPC_LAUNCH(cell_nn_kernel<uint32_t>, blocks, kThreads, 0, as_stream(stream), make_list<uint32_t>(grid_host, cell_keys, ncells, start, pos_sorted, xyzw, n),
                qx, qy, qz, (long long)nq, bounded != 0, nn_idx, nn_dist);
  } else {
    VAPC_LAUNCH(cell_nn_kernel<unsigned long long>, blocks, kThreads, 0, as_stream(stream),
                make_list<unsigned long long>(grid_host, cell_keys, ncells, start, pos_sorted, xyzw, n), qx, qy, qz, (long long)nq, bounded != 0, nn_idx,
                nn_dist);
  }
  VAPC_CHECK_LAUNCH();
  return VAPC_OK;
}

extern "C" int vapc_nn_all_pairs(const double* qx, const double* qy, const double* qz, const int64_t* which, int64_t nwhich, const double* xyzw,
                                 int64_t n, uint32_t* nn_idx, double* nn_dist, void* stream) {
  if (nwhich < 0 || n <= 0) return vapc_set_error(VAPC_E_INVALID, "vapc_nn_all_pairs: bad sizes");
  if (nwhich == 0) return VAPC_OK;
  if (!qx || !qy || !qz || !which || !xyzw || !nn_idx || !nn_dist) return vapc_set_error(VAPC_E_INVALID, "vapc_nn_all_pairs: null argument");
  VAPC_LAUNCH(nn_brute_kernel, (unsigned)((nwhich + kThreads - 1) / kThreads), kThreads, 0, as_stream(stream), reinterpret_cast<const double4*>(xyzw),
              (long long)n, qx, qy, qz, reinterpret_cast<const long long*>(which), (long long)nwhich, nn_idx, nn_dist);
  VAPC_CHECK_LAUNCH();
  return VAPC_OK;
}

extern "C" int vapc_selftest_cell_nn_host(const double* qx, const double* qy, const double* qz, int64_t nq, const vapc_grid_t* grid_host,
                                          const void* cell_keys, int64_t ncells, const uint32_t* start, const uint32_t* pos_sorted,
                                          const double* xyzw, int64_t n, int32_t bounded, uint32_t* nn_idx, double* nn_dist) {
  if (nq < 0) return vapc_set_error(VAPC_E_INVALID, "vapc_selftest_cell_nn_host: negative size");
  if (const int rc = check_list("vapc_selftest_cell_nn_host", grid_host, cell_keys, ncells, start, pos_sorted, xyzw, n)) return rc;
  if (grid_host->key_bytes == 4) {
    const auto cl = make_list<uint32_t>(grid_host, cell_keys, ncells, start, pos_sorted, xyzw, n);
    for (int64_t i = 0; i < nq; ++i) cell_nn_one(cl, qx[i], qy[i], qz[i], bounded != 0, nn_idx + i, nn_dist + i);
  } else {
    const auto cl = make_list<unsigned long long>(grid_host, cell_keys, ncells, start, pos_sorted, xyzw, n);
    for (int64_t i = 0; i < nq; ++i) cell_nn_one(cl, qx[i], qy[i], qz[i], bounded != 0, nn_idx + i, nn_dist + i);
  }
  return VAPC_OK;
}

extern "C" size_t vapc_cluster_workspace_bytes(int64_t n) { return (size_t)(n > 0 ? n : 0) * 8 + 256; }

extern "C" int vapc_cluster_components(const vapc_grid_t* grid_host, const void* cell_keys, int64_t ncells, const uint32_t* start,
                                       const uint32_t* pos_sorted, const double* xyzw, int64_t n, double distance, const uint32_t* first,
                                       const uint32_t* weight, uint32_t* root, uint32_t* hop2_first, uint32_t* comp_first,
                                       int64_t* comp_size, void* ws, size_t ws_bytes, void* stream) {
  if (const int rc = check_list("vapc_cluster_components", grid_host, cell_keys, ncells, start, pos_sorted, xyzw, n)) return rc;
  if (!root || !hop2_first || !comp_first || !comp_size || !ws) return vapc_set_error(VAPC_E_INVALID, "vapc_cluster_components: null argument");
  if (ws_bytes < vapc_cluster_workspace_bytes(n)) return vapc_set_error(VAPC_E_WORKSPACE, "vapc_cluster_components: workspace too small");
  if (!(distance <= grid_host->voxel_size)) return vapc_set_error(VAPC_E_INVALID, "vapc_cluster_components: the cell size must be >= the cluster distance");
  uint32_t* parent = reinterpret_cast<uint32_t*>(ws);
  uint32_t* m1 = parent + n;
  cudaStream_t s = as_stream(stream);
  unsigned long long* sizes = reinterpret_cast<unsigned long long*>(comp_size);
  VAPC_LAUNCH(cluster_init_kernel, (unsigned)((n + 255) / 256), 256, 0, s, parent, comp_first, sizes, (long long)n);
  VAPC_CHECK_LAUNCH();
  const unsigned blocks = (unsigned)((n + kThreads - 1) / kThreads);
  if (grid_host->key_bytes == 4) {
    const auto cl = make_list<uint32_t>(grid_host, cell_keys, ncells, start, pos_sorted, xyzw, n);
    VAPC_LAUNCH(cluster_link_kernel<uint32_t>, blocks, kThreads, 0, s, cl, distance, first, parent, m1);
    VAPC_CHECK_LAUNCH();
    VAPC_LAUNCH(cluster_hop2_kernel<uint32_t>, blocks, kThreads, 0, s, cl, distance, m1, hop2_first);
  } else {
    const auto cl = make_list<unsigned long long>(grid_host, cell_keys, ncells, start, pos_sorted, xyzw, n);
    VAPC_LAUNCH(cluster_link_kernel<unsigned long long>, blocks, kThreads, 0, s, cl, distance, first, parent, m1);
    VAPC_CHECK_LAUNCH();
    VAPC_LAUNCH(cluster_hop2_kernel<unsigned long long>, blocks, kThreads, 0, s, cl, distance, m1, hop2_first);
  }
  VAPC_CHECK_LAUNCH();
  VAPC_LAUNCH(cluster_finish_kernel, (unsigned)((n + 255) / 256), 256, 0, s, parent, first, weight, root, comp_first, sizes, (long long)n);
  VAPC_CHECK_LAUNCH();
  return VAPC_OK;
}

extern "C" int vapc_selftest_cluster_host(const vapc_grid_t* grid_host, const void* cell_keys, int64_t ncells, const uint32_t* start,
                                          const uint32_t* pos_sorted, const double* xyzw, int64_t n, double distance, const uint32_t* first,
                                          const uint32_t* weight, uint32_t* root, uint32_t* hop2_first, uint32_t* comp_first,
                                          int64_t* comp_size) {
  if (const int rc = check_list("vapc_selftest_cluster_host", grid_host, cell_keys, ncells, start, pos_sorted, xyzw, n)) return rc;
  if (!(distance <= grid_host->voxel_size)) return vapc_set_error(VAPC_E_INVALID, "vapc_selftest_cluster_host: the cell size must be >= the cluster distance");
  uint32_t* scratch = new uint32_t[2 * (size_t)n];
  if (grid_host->key_bytes == 4)
    cluster_host(make_list<uint32_t>(grid_host, cell_keys, ncells, start, pos_sorted, xyzw, n), distance, first, weight, root, hop2_first, comp_first,
                 comp_size, scratch, scratch + n);
  else
    cluster_host(make_list<unsigned long long>(grid_host, cell_keys, ncells, start, pos_sorted, xyzw, n), distance, first, weight, root, hop2_first,
                 comp_first, comp_size, scratch, scratch + n);
  delete[] scratch;
  return VAPC_OK;
}

extern "C" int vapc_reduce_candidates(const int64_t* rows, const double* dist, const int64_t* idx, int64_t n, int64_t nrows, double* best_dist,
                                      int64_t* best_idx, void* stream) {
  if (n < 0 || nrows < 0) return vapc_set_error(VAPC_E_INVALID, "vapc_reduce_candidates: negative size");
  if (nrows == 0) return VAPC_OK;
  if (!best_dist || !best_idx || (n > 0 && (!rows || !dist || !idx))) return vapc_set_error(VAPC_E_INVALID, "vapc_reduce_candidates: null buffer");
  cudaStream_t s = as_stream(stream);
  unsigned long long* bd = reinterpret_cast<unsigned long long*>(best_dist);
  unsigned long long* bi = reinterpret_cast<unsigned long long*>(best_idx);
  VAPC_LAUNCH(cand_init_kernel, (unsigned)((nrows + 255) / 256), 256, 0, s, bd, bi, (long long)nrows);
  VAPC_CHECK_LAUNCH();
  if (n == 0) return VAPC_OK;
  const unsigned blocks = (unsigned)((n + 255) / 256);
  VAPC_LAUNCH(cand_min_dist_kernel, blocks, 256, 0, s, reinterpret_cast<const long long*>(rows), dist, (long long)n, bd);
  VAPC_CHECK_LAUNCH();
  VAPC_LAUNCH(cand_min_idx_kernel, blocks, 256, 0, s, reinterpret_cast<const long long*>(rows), dist, reinterpret_cast<const long long*>(idx), (long long)n, bd, bi);
  VAPC_CHECK_LAUNCH();
  return VAPC_OK;
}
