// Layout of the segmentation workspace (vapc_segment_workspace_bytes) shared by the kernels that fill it
// (count_heads_kernel in segment_reduce.cu, the cell sort in cell_sort.cu) and the tile reductions that read it.
#pragma once
#include "common.cuh"

namespace segws {
constexpr int kSegTile = 1024;  // sorted positions per tile of the segmented reductions
struct SegWs {  // layout of the segmentation workspace
  uint32_t* tile_heads;       // [ntiles]   heads per tile, then (in place) exclusive scan = first voxel row of the tile
  unsigned long long* total;  // [1]        number of voxels
  void* scan_ws;              // scan scratch
  uint32_t* carry_cnt;        // [ntiles]   points in the leading partial segment (0 = none)
  double* carry;              // [ntiles][10] its partial moments (or other per-op payload)
  uint32_t* long_chains;      // [ntiles]   first tile of every carry chain longer than one tile
  uint32_t* n_long;           // [1]
};
inline int64_t seg_tiles(int64_t n) { return (n + kSegTile - 1) / kSegTile; }
inline size_t align_up(size_t v, size_t a = 256) { return (v + a - 1) / a * a; }
inline size_t seg_ws_bytes(int64_t n) {
  const int64_t nt = seg_tiles(n < 1 ? 1 : n);
  return align_up(nt * 4) + align_up(8) + align_up(scan_u32_workspace_bytes(nt)) + align_up(nt * 4) + align_up(nt * 10 * 8) +
         align_up(nt * 4) + align_up(4);
}
inline SegWs carve(void* ws, int64_t n) {
  const int64_t nt = seg_tiles(n < 1 ? 1 : n);
  unsigned char* p = static_cast<unsigned char*>(ws);
  SegWs w;
  w.tile_heads = reinterpret_cast<uint32_t*>(p); p += align_up(nt * 4);
  w.total = reinterpret_cast<unsigned long long*>(p); p += align_up(8);
  w.scan_ws = p; p += align_up(scan_u32_workspace_bytes(nt));
  w.carry_cnt = reinterpret_cast<uint32_t*>(p); p += align_up(nt * 4);
  w.carry = reinterpret_cast<double*>(p); p += align_up(nt * 10 * 8);
  w.long_chains = reinterpret_cast<uint32_t*>(p); p += align_up(nt * 4);
  w.n_long = reinterpret_cast<uint32_t*>(p);
  return w;
}

}  // namespace segws
