"""Device-level pipeline over libvapc_b200: torch tensors are only the containers (device
memory, streams); every computation is a call through the C ABI (include/vapc_b200.h).

    VoxelCloud.build(x, y, z, voxel_size, origin)   keys -> radix sort -> segmentation -> moments
    cloud.stats(...)                                density / centroid / std / cov / eig / features
    cloud.closest_to_cog(), cloud.mode_count(...), mask_membership(...), join_epochs(...), ...

This is what `vapc_b200.Vapc` / `BiTemporalVapc` (the reference-shaped facade) drive.
"""
from __future__ import annotations

import ctypes as C
import os
from dataclasses import dataclass
from typing import Dict, Optional, Sequence

import numpy as np
import torch

from . import _lib as L
from ._lib import Grid, LasTransform

FEATURE_NAMES = (
    "Sum_of_Eigenvalues", "Omnivariance", "Eigenentropy", "Anisotropy", "Planarity", "Linearity",
    "Surface_Variation", "Sphericity",
)


def _require_cuda():
    if not torch.cuda.is_available():
        raise RuntimeError("vapc_b200 needs a CUDA device (B200, sm_100a); there is no CPU fallback")


def _ptr(t: Optional[torch.Tensor]):
    return C.c_void_p(0) if t is None else C.c_void_p(t.data_ptr())


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _dev_f64(a, device) -> torch.Tensor:
    """A contiguous fp64 CUDA tensor from a tensor / numpy array / pandas column."""
    if isinstance(a, torch.Tensor):
        t = a
    else:
        t = torch.from_numpy(np.ascontiguousarray(np.asarray(a, dtype=np.float64)))
    return t.to(device=device, dtype=torch.float64, non_blocking=True).contiguous()


def _u32_t(t: torch.Tensor) -> torch.Tensor:
    """uint32 values stored in an int32 tensor -> int64."""
    return t.to(torch.int64) & 0xFFFFFFFF


def _empty(n, dtype, device):
    return torch.empty(int(n), dtype=dtype, device=device)


def make_grid(voxel_size: float, origin: Sequence[float]) -> Grid:
    if not voxel_size > 0:
        raise ValueError("voxel_size must be a positive number.")
    if len(origin) != 3:
        raise ValueError("origin must be a list of three coordinates [x, y, z].")
    g = Grid()
    g.voxel_size = float(voxel_size)
    for k in range(3):
        g.origin[k] = float(origin[k])
    return g


def grid_from_bounds(voxel_size, origin, minmax6) -> Grid:
    g = make_grid(voxel_size, origin)
    mm = (C.c_double * 6)(*[float(v) for v in minmax6])
    L.call("vapc_grid_from_bounds", C.byref(g), mm)
    return g


def grid_from_voxel_bounds(voxel_size, origin, vmin, vmax) -> Grid:
    """Grid covering integer voxel bounds [vmin, vmax] per axis (used for masks / joins)."""
    g = make_grid(voxel_size, origin)
    total = 0
    for k in range(3):
        g.vmin[k] = int(vmin[k])
        g.bits[k] = max(0, int(int(vmax[k]) - int(vmin[k])).bit_length())
        total += g.bits[k]
    if total > 64:
        raise L.VapcError(L.VAPC_E_KEYSPACE, f"voxel bounding box needs {total} key bits (> 64)")
    g.key_bytes = 4 if total <= 32 else 8
    return g


def coordinate_bounds(x, y, z) -> np.ndarray:
    """[min x, min y, min z, max x, max y, max z] of device columns (one tiny D2H copy)."""
    n = x.numel()
    out = _empty(6, torch.float64, x.device)
    ws_bytes = L.raw("vapc_minmax_workspace_bytes")()
    ws = _empty(ws_bytes, torch.uint8, x.device)
    L.call("vapc_minmax_f64", _ptr(x), _ptr(y), _ptr(z), n, _ptr(out), _ptr(ws), ws_bytes, _stream())
    return out.cpu().numpy()


def las_transform(scales: Sequence[float], offsets: Sequence[float]) -> LasTransform:
    t = LasTransform()
    for k in range(3):
        t.scale[k] = float(scales[k])
        t.offset[k] = float(offsets[k])
    return t


def _dev_i32(a, device) -> torch.Tensor:
    """A contiguous int32 CUDA tensor from a tensor / numpy array (LAS integer coordinates)."""
    if not isinstance(a, torch.Tensor):
        a = np.asarray(a)
        if a.dtype != np.int32:
            raise TypeError("LAS integer coordinates must be int32")
        a = torch.from_numpy(np.ascontiguousarray(a))
    t = a
    if t.dtype != torch.int32:
        raise TypeError("LAS integer coordinates must be int32")
    return t.to(device=device, non_blocking=True).contiguous()


def coordinate_bounds_i32(X, Y, Z, t: LasTransform) -> np.ndarray:
    """coordinate_bounds of the doubles X * scale + offset that int32 LAS coordinates stand for (12 B/point read)."""
    out = _empty(6, torch.float64, X.device)
    ws_bytes = L.raw("vapc_minmax_workspace_bytes")()
    ws = _empty(ws_bytes, torch.uint8, X.device)
    L.call("vapc_minmax_i32", _ptr(X), _ptr(Y), _ptr(Z), X.numel(), C.byref(t), _ptr(out), _ptr(ws), ws_bytes, _stream())
    return out.cpu().numpy()


def las_coordinates(X, Y, Z, scales, offsets, device=None):
    """The fp64 x, y, z columns of int32 LAS coordinates, X * scale + offset (what laspy hands the reference's DataHandler,
    datahandler.py:72-95), computed on the device; bit-identical to numpy's X * scale + offset."""
    _require_cuda()
    if device is None:
        device = X.device if isinstance(X, torch.Tensor) and X.is_cuda else torch.device("cuda", torch.cuda.current_device())
    X, Y, Z = (_dev_i32(a, device) for a in (X, Y, Z))
    t = las_transform(scales, offsets)
    out = [_empty(X.numel(), torch.float64, device) for _ in range(3)]
    L.call("vapc_las_coords_i32", _ptr(X), _ptr(Y), _ptr(Z), X.numel(), C.byref(t), _ptr(out[0]), _ptr(out[1]), _ptr(out[2]), _stream())
    return out


def voxel_coords(x, y, z, voxel_size, origin):
    """Vapc.voxelize (vapc.py:188-207): three int64 device columns, bit-exact with numpy."""
    _require_cuda()
    dev = x.device if isinstance(x, torch.Tensor) and x.is_cuda else torch.device("cuda", torch.cuda.current_device())
    x, y, z = (_dev_f64(a, dev) for a in (x, y, z))
    g = make_grid(voxel_size, origin)
    n = x.numel()
    out = [_empty(n, torch.int64, dev) for _ in range(3)]
    L.call("vapc_voxel_coords_f64", _ptr(x), _ptr(y), _ptr(z), n, C.byref(g), _ptr(out[0]), _ptr(out[1]), _ptr(out[2]), _stream())
    return out


def sort_pairs(keys: torch.Tensor, end_bit: int, idx: Optional[torch.Tensor] = None, preserve_keys: bool = False, begin_bit: int = 0):
    """Stable LSD radix sort of (key, idx) by the low `end_bit` key bits.  keys: uint32-as-int32 or
    uint64-as-int64 storage (torch has no unsigned 32/64 arithmetic; only the bytes matter).
    Returns (keys_sorted, idx_sorted) — new tensors or the inputs, whichever holds the result.
    preserve_keys: sort through a third buffer so that `keys` keeps its contents."""
    n = keys.numel()
    dev = keys.device
    key_bytes = keys.element_size()
    keys_alt = torch.empty_like(keys)
    passes = (int(end_bit) - int(begin_bit) + 7) // 8
    keys_alt2 = torch.empty_like(keys) if (preserve_keys and passes > 1) else None
    iota = idx is None
    if iota:
        idx = _empty(n, torch.int32, dev)
    idx_alt = torch.empty_like(idx)
    ws_bytes = L.raw("vapc_sort_workspace_bytes")(n)
    ws = _empty(ws_bytes, torch.uint8, dev)
    loc = C.c_int32(0)
    L.call("vapc_sort_pairs_bits", _ptr(keys), _ptr(keys_alt), _ptr(keys_alt2), _ptr(idx), _ptr(idx_alt), 1 if iota else 0, n, key_bytes,
           int(begin_bit), int(end_bit), _ptr(ws), ws_bytes, C.byref(loc), _stream())
    out_keys = (keys, keys_alt, keys_alt2)[loc.value & 3]
    if preserve_keys and out_keys is keys and passes > 0:  # cannot happen with a third buffer; guard the contract
        raise AssertionError("sort overwrote preserved keys")
    return out_keys, (idx_alt if loc.value & 4 else idx)


def sort_pairs_segmented(keys: torch.Tensor, end_bit: int, seg_start: torch.Tensor, nseg: int, idx: Optional[torch.Tensor] = None):
    """Stable sort of every segment [seg_start[s], seg_start[s+1]) by the low `end_bit` key bits, all segments in
    the same launches (`keys` is overwritten).  Returns (keys_sorted, idx_sorted); idx defaults to the global
    position of every element."""
    n = keys.numel()
    dev = keys.device
    keys_alt = torch.empty_like(keys)
    iota = idx is None
    if iota:
        idx = _empty(n, torch.int32, dev)
    idx_alt = torch.empty_like(idx)
    ws_bytes = L.raw("vapc_sort_workspace_bytes")(n)
    ws = _empty(ws_bytes, torch.uint8, dev)
    loc = C.c_int32(0)
    L.call("vapc_sort_pairs_segmented", _ptr(keys), _ptr(keys_alt), _ptr(idx), _ptr(idx_alt), 1 if iota else 0, n, keys.element_size(), int(end_bit),
           _ptr(seg_start), int(nseg), _ptr(ws), ws_bytes, C.byref(loc), _stream())
    return (keys, keys_alt)[loc.value & 3], (idx_alt if loc.value & 4 else idx)


_L2_FETCH_SET = set()


def _tune_device(device):
    """Once per device: ask the L2 for 32-byte fetch granularity (random 32-byte record gathers)."""
    import os

    i = device.index if device.index is not None else torch.cuda.current_device()
    if i in _L2_FETCH_SET:
        return
    _L2_FETCH_SET.add(i)
    g = int(os.environ.get("VAPC_L2_FETCH", "32"))
    if g:
        with torch.cuda.device(i):
            L.call("vapc_set_l2_fetch_granularity", g)


CELL_SORT = os.environ.get("VAPC_CELL_SORT", "0") != "0"  # dense grids: counting sort through a per-bucket cell table
CELL_SORT_MIN_BITS = 9   # fewer key bits below the bucket digit: a single radix pass is as cheap
CELL_SORT_MAX_MEAN = 16  # mean points per cell of the key space beyond which the per-cell ordering would dominate
P2V_SIDE_STREAM = os.environ.get("VAPC_P2V_SIDE_STREAM", "1") != "0"  # point->voxel lookups beside the next kernel of the caller's stream
P2V_SIDE_CTAS = int(os.environ.get("VAPC_P2V_SIDE_CTAS", "4"))             # their CTAs per SM there (alone: 16)
P2V_SIDE_PRIORITY = os.environ.get("VAPC_P2V_SIDE_PRIORITY", "1") != "0"
_SIDE_STREAMS: Dict[int, "torch.cuda.Stream"] = {}


_PINNED_SCALARS: Dict[int, tuple] = {}


def _pinned_scalars(device):
    """(int64[2], int32[1]) pinned host words per device for the voxel count / status read-back of build()."""
    i = device.index if device.index is not None else torch.cuda.current_device()
    if i not in _PINNED_SCALARS:
        _PINNED_SCALARS[i] = (torch.zeros(2, dtype=torch.int64).pin_memory(), torch.zeros(1, dtype=torch.int32).pin_memory())
    return _PINNED_SCALARS[i]


def _side_stream(device):
    i = device.index if device.index is not None else torch.cuda.current_device()
    if i not in _SIDE_STREAMS:
        # high priority: its CTAs are placed as soon as the kernel in front on the caller's stream frees a slot, not behind that kernel's queue
        _SIDE_STREAMS[i] = torch.cuda.Stream(device=i, priority=-1 if P2V_SIDE_PRIORITY else 0)
    return _SIDE_STREAMS[i]


DENSE_P2V_MAX_BITS = 29  # rank-bitmap point->voxel table: 2^bits / 4 bytes, <= 128 MB stays L2-resident on B200


@dataclass
class VoxelStats:
    density: Optional[torch.Tensor] = None  # [V]
    cog: Optional[torch.Tensor] = None  # [3, V]
    std: Optional[torch.Tensor] = None  # [3, V]
    cov: Optional[torch.Tensor] = None  # [6, V]  xx xy xz yy yz zz
    eig: Optional[torch.Tensor] = None  # [3, V]
    eig_n: Optional[torch.Tensor] = None  # [3, V]
    feat: Optional[torch.Tensor] = None  # [8, V]  FEATURE_NAMES order


class VoxelCloud:
    """A point cloud voxelised on the GPU: sorted (key, index) pairs, the voxel table in
    lexicographic (vx, vy, vz) order, the (count, mean, M2) moments and the point->voxel map."""

    def __init__(self):
        self.grid: Grid = None
        self.n = 0
        self.nvoxels = 0
        self.x = self.y = self.z = None
        self.xyzw = None  # 32-byte records {x, y, z, original index}, in bucket order (or input order)
        self.keys_sorted = None
        self.pos_sorted = None  # position of each sorted point's record in xyzw
        self._idx_sorted = None  # original index of each sorted point (formed on first use)
        self.voxel_keys = None
        self.start = None
        self.count = None
        self.first_idx = None
        self.mean3 = None
        self.m2 = None
        self._point2voxel = None
        self._p2v_event = None  # set while the point->voxel map is still being formed on the side stream
        self._p2v_hold = None   # (side stream, tensors the lookups there still read)
        self._seg_ws = None
        self.sorted_by = "radix"  # or "cells" (vapc_sort_cells_segmented)
        self._stats_cache: Dict[str, torch.Tensor] = {}
        self._centroid_formula = "exact"
        self.axis_order = (0, 1, 2)  # coordinate axis behind every internal axis of the keys / moments (multi-GPU tables permute them)

    # ------------------------------------------------------------------------------------
    @classmethod
    def build(cls, x, y, z, voxel_size, origin=(0.0, 0.0, 0.0), *, bounds=None, want_point2voxel=True,
              keep_columns=True, device=None, bucketed=True, las=None, cell_sort=None, voxels=None) -> "VoxelCloud":
        """voxelize + groupby of the reference (vapc.py:188-207 and the groupby at :724, :843, :959)
        as: bounds -> packed keys + xyzw records in key-prefix bucket order -> stable radix sort of the remaining key
        bits inside the buckets -> segmentation -> moments.  bucketed=False keeps the records in input order and
        sorts all key bits (same results; the per-voxel gathers then miss the L2).
        las = (X, Y, Z, scales, offsets): the cloud as the int32 coordinates of a LAS file (x, y, z are then ignored and may be
        None): the key kernels read 12 B/point and form X * scale + offset themselves — the same doubles, the same results.
        voxels = (vx, vy, vz): the voxel triple of every point GIVEN by the caller (int64 columns) instead of derived from the
        coordinates — the reference groups by the frame's voxel_x / voxel_y / voxel_z columns whatever they hold (vapc.py:724, :843,
        :959), e.g. after the caller edited them or the coordinates.  Moments stay relative to c(v) of the given voxel."""
        _require_cuda()
        if voxels is not None:
            bucketed = False
            if las is not None:
                raise ValueError("given voxel triples and LAS integer input exclude each other")
        if cell_sort is None:
            cell_sort = CELL_SORT
        if las is not None:
            if not bucketed:
                raise ValueError("LAS integer input needs the bucketed path")
            X0 = las[0]
            if device is None:
                device = X0.device if isinstance(X0, torch.Tensor) and X0.is_cuda else torch.device("cuda", torch.cuda.current_device())
            lX, lY, lZ = (_dev_i32(a, device) for a in las[:3])
            lt = las_transform(las[3], las[4])
            n = lX.numel()
            if lY.numel() != n or lZ.numel() != n:
                raise ValueError("X, Y, Z must have the same length")
            if keep_columns:
                x, y, z = las_coordinates(lX, lY, lZ, las[3], las[4], device)
            if bounds is None and n:
                bounds = coordinate_bounds_i32(lX, lY, lZ, lt)
        else:
            if device is None:
                device = x.device if isinstance(x, torch.Tensor) and x.is_cuda else torch.device("cuda", torch.cuda.current_device())
            x, y, z = (_dev_f64(a, device) for a in (x, y, z))
            n = x.numel()
            if y.numel() != n or z.numel() != n:
                raise ValueError("X, Y, Z must have the same length")
        if n == 0:
            raise ValueError("empty point cloud")
        if n >= 2 ** 32:
            raise ValueError("at most 2**32 - 1 points per GPU")
        self = cls()
        self.n = n
        if keep_columns:
            self.x, self.y, self.z = x, y, z
        if bounds is None:
            bounds = coordinate_bounds(x, y, z)
        self.grid = g = grid_from_bounds(voxel_size, origin, bounds)
        if voxels is not None:
            gv = [torch.as_tensor(np.asarray(a, dtype=np.int64) if not isinstance(a, torch.Tensor) else a).to(device=device, dtype=torch.int64).contiguous()
                  for a in voxels]
            if any(a.numel() != n for a in gv):
                raise ValueError("voxel columns must have the length of the cloud")
            coord_grid = g  # records are written with the coordinates' own grid; the keys below replace its keys
            self.grid = g = grid_from_voxel_bounds(voxel_size, origin, [int(a.min().item()) for a in gv], [int(a.max().item()) for a in gv])
            g.key_bytes = 8  # vapc_pack_triples_i64 writes 64-bit keys
            if g.total_bits > 63:
                raise L.VapcError(L.VAPC_E_KEYSPACE, "given voxel triples need more than 63 key bits")
        kdtype = torch.int32 if g.key_bytes == 4 else torch.int64
        self.xyzw = torch.empty((n, 4), dtype=torch.float64, device=device)
        status = torch.zeros(1, dtype=torch.int32, device=device)
        _tune_device(device)
        dense_p2v = want_point2voxel and g.total_bits <= DENSE_P2V_MAX_BITS
        ws_bytes = L.raw("vapc_segment_workspace_bytes")(n)
        self._seg_ws = _empty(ws_bytes, torch.uint8, device)
        nv = torch.zeros(2, dtype=torch.int64, device=device)
        host = None
        if bucketed:
            keys = _empty(n, kdtype, device)  # keys in input order
            keys_b = _empty(n, kdtype, device)
            bucket_start = _empty(257, torch.int32, device)
            bws_bytes = L.raw("vapc_bucket_workspace_bytes")(n)
            bws = _empty(bws_bytes, torch.uint8, device)
            bucket_bits, bkey_bytes = C.c_int32(0), C.c_int32(0)
            if las is not None:
                L.call("vapc_pack_records_bucketed_i32", _ptr(lX), _ptr(lY), _ptr(lZ), n, C.byref(lt), C.byref(g), _ptr(keys_b), _ptr(self.xyzw), _ptr(keys),
                       _ptr(bucket_start), C.byref(bucket_bits), C.byref(bkey_bytes), _ptr(status), _ptr(bws), bws_bytes, _stream())
            else:
                L.call("vapc_pack_records_bucketed", _ptr(x), _ptr(y), _ptr(z), n, C.byref(g), _ptr(keys_b), _ptr(self.xyzw), _ptr(keys), _ptr(bucket_start),
                       C.byref(bucket_bits), C.byref(bkey_bytes), _ptr(status), _ptr(bws), bws_bytes, _stream())
            del bws
            low_bits = g.total_bits - bucket_bits.value
            nbuckets = 1 << bucket_bits.value
            # dense grids: the cells of a bucket fit a shared-memory table -> counting sort instead of radix passes (csrc/cell_sort.cu)
            if (cell_sort and g.key_bytes == 4 and bucket_bits.value == 8 and CELL_SORT_MIN_BITS <= low_bits <= 16
                    and n <= (CELL_SORT_MAX_MEAN << g.total_bits)):
                self.keys_sorted = _empty(n, kdtype, device)
                self.pos_sorted = _empty(n, torch.int32, device)
                L.call("vapc_sort_cells_segmented", _ptr(keys_b), n, 4, low_bits, _ptr(bucket_start), nbuckets, _ptr(self.keys_sorted),
                       _ptr(self.pos_sorted), _ptr(nv), _ptr(status), _ptr(self._seg_ws), ws_bytes, _stream())
                host = torch.cat([nv[:1], status.to(torch.int64)]).cpu()  # the one sync of the pipeline
                if int(host[1]) & 4:  # a cell with too many points for the per-cell ordering: the radix sort takes over
                    host = None
                    self.keys_sorted = self.pos_sorted = None
                else:
                    self.sorted_by = "cells"
            if host is None:
                if bkey_bytes.value == 4 and g.key_bytes == 8:
                    # 64-bit grid, but the bits below the bucket digit fit 32: sort 4-byte keys, then rebuild the full sorted keys
                    low_sorted, self.pos_sorted = sort_pairs_segmented(keys_b.view(torch.int32)[:n], low_bits, bucket_start, nbuckets)
                    self.keys_sorted = _empty(n, torch.int64, device)
                    L.call("vapc_expand_bucket_keys", _ptr(low_sorted), n, _ptr(bucket_start), bucket_bits.value, low_bits, _ptr(self.keys_sorted), _stream())
                    del low_sorted
                else:
                    self.keys_sorted, self.pos_sorted = sort_pairs_segmented(keys_b, low_bits, bucket_start, nbuckets)
            del keys_b
            if not dense_p2v:
                del keys
        else:
            keys = _empty(n, kdtype, device)
            if voxels is not None:
                scratch = _empty(n, torch.int32 if coord_grid.key_bytes == 4 else torch.int64, device)
                L.call("vapc_pack_keys_f64", _ptr(x), _ptr(y), _ptr(z), n, C.byref(coord_grid), _ptr(scratch), _ptr(self.xyzw), _ptr(status), _stream())
                del scratch
                L.call("vapc_pack_triples_i64", _ptr(gv[0]), _ptr(gv[1]), _ptr(gv[2]), n, C.byref(g), _ptr(keys), _ptr(status), _stream())
                del gv
            else:
                L.call("vapc_pack_keys_f64", _ptr(x), _ptr(y), _ptr(z), n, C.byref(g), _ptr(keys), _ptr(self.xyzw), _ptr(status), _stream())
            self.keys_sorted, self.pos_sorted = sort_pairs(keys, g.total_bits, preserve_keys=dense_p2v)
            if not dense_p2v:
                del keys
        if host is None:
            L.call("vapc_count_voxels", _ptr(self.keys_sorted), n, g.key_bytes, _ptr(nv), _ptr(self._seg_ws), ws_bytes, _stream())
            # the one sync of the pipeline: two small copies into pinned memory, no kernel in between (the GPU idles until the host
            # has read V and queued the next call, so every microsecond here is on the step)
            hv, hs = _pinned_scalars(device)
            hv.copy_(nv, non_blocking=True)
            hs.copy_(status, non_blocking=True)
            torch.cuda.current_stream(device).synchronize()
            host = (int(hv[0]), int(hs[0]))
        if int(host[1]) & 1:
            raise L.VapcError(L.VAPC_E_RANGE, "a point fell outside the voxel grid bounds (NaN / inf coordinate or wrong bounds)")
        self.nvoxels = V = int(host[0])
        self.voxel_keys = _empty(V, kdtype, device)
        self.start = _empty(V + 1, torch.int32, device)
        self.count = _empty(V, torch.int32, device)
        self.first_idx = _empty(V, torch.int32, device)
        self.mean3 = torch.empty((3, V), dtype=torch.float64, device=device)
        self.m2 = torch.empty((6, V), dtype=torch.float64, device=device)
        self.point2voxel = _empty(n, torch.int32, device) if want_point2voxel else None
        row_sorted = _empty(n, torch.int32, device) if (want_point2voxel and not dense_p2v) else None
        if row_sorted is not None:
            self._idx_sorted = _empty(n, torch.int32, device)
        L.call("vapc_reduce_moments", _ptr(self.keys_sorted), _ptr(self.pos_sorted), n, C.byref(g), _ptr(self.xyzw), V,
               _ptr(self.voxel_keys), _ptr(self.start), _ptr(self.count), _ptr(self.first_idx), _ptr(self.mean3), _ptr(self.m2),
               _ptr(row_sorted), _ptr(self._idx_sorted), _ptr(self._seg_ws), ws_bytes, _stream())
        if row_sorted is not None:
            # (index, row) pairs grouped by the leading byte of the index, then scattered inside L2-sized windows
            nbits = max(1, int(n - 1).bit_length())
            grouped_idx, grouped_row = sort_pairs(self._idx_sorted, nbits, idx=row_sorted, preserve_keys=True, begin_bit=max(0, nbits - 8))
            L.call("vapc_scatter_u32", _ptr(grouped_idx), _ptr(grouped_row), n, _ptr(self.point2voxel), _stream())
            del grouped_idx, grouped_row, row_sorted
        if dense_p2v:
            tbytes = L.raw("vapc_point2voxel_table_bytes")(g.total_bits)
            if P2V_SIDE_STREAM:
                # The lookups are bound by the rate of uncoalesced gathers (few issue slots, little DRAM traffic); whatever the caller
                # launches next on its own stream — normally vapc_voxel_stats, bound by issue slots and the FP64 pipe — shares the SMs
                # with them.  Reading .point2voxel joins the streams.
                cur, side = torch.cuda.current_stream(device), _side_stream(device)
                side.wait_stream(cur)
                with torch.cuda.stream(side):
                    table = _empty(tbytes, torch.uint8, device)
                    L.call("vapc_point2voxel_set_ctas_per_sm", P2V_SIDE_CTAS)
                    L.call("vapc_point2voxel_dense", _ptr(keys), n, _ptr(self.voxel_keys), V, g.key_bytes, g.total_bits, _ptr(table), tbytes,
                           _ptr(self._point2voxel), _stream())
                    L.call("vapc_point2voxel_set_ctas_per_sm", 16)
                    self._p2v_event = side.record_event()
                # The lookups read `keys` and the voxel keys and write the map on the side stream.  They are HELD here until the caller's
                # stream has waited for the event (the first read of .point2voxel) instead of being handed to the allocator with
                # record_stream: a block whose release waits for a side-stream event is not back in the pool when the next build asks for
                # the same size a few microseconds later, and the cudaMalloc that follows stalls that step by 3-4 ms every now and then
                # (bench step_ms).  Only a cloud that is dropped before anybody read the map falls back to record_stream (__del__).
                self._p2v_hold = (side, keys)
            else:
                table = _empty(tbytes, torch.uint8, device)
                L.call("vapc_point2voxel_dense", _ptr(keys), n, _ptr(self.voxel_keys), V, g.key_bytes, g.total_bits, _ptr(table), tbytes,
                       _ptr(self._point2voxel), _stream())
            del keys, table
        return self

    @property
    def point2voxel(self):
        """uint32 voxel row of every point (int32 storage).  Formed on a side stream by build(); the first read makes the caller's
        stream wait for it."""
        if self._p2v_event is not None:
            torch.cuda.current_stream(self._point2voxel.device).wait_event(self._p2v_event)
            self._p2v_event = None
            self._p2v_hold = None
        return self._point2voxel

    @point2voxel.setter
    def point2voxel(self, t):
        self._p2v_event = None
        self._point2voxel = t

    def __del__(self):
        # dropped while the side stream may still be reading / writing: the allocator must not reuse these blocks before it is done
        hold = getattr(self, "_p2v_hold", None)
        if hold is not None and getattr(self, "_p2v_event", None) is not None:
            try:
                for t in (hold[1], self.voxel_keys, self._point2voxel):
                    if t is not None:
                        t.record_stream(hold[0])
            except Exception:  # noqa: BLE001  (interpreter shutdown)
                pass

    @property
    def device(self):
        return self.voxel_keys.device

    @property
    def idx_sorted(self):
        """Original index of every sorted position (int32 storage of uint32).  Only the per-attribute statistics and the grouped
        point->voxel scatter read it, so it is formed on first use from the records (4 B/point written) instead of by every build."""
        if self._idx_sorted is None:
            self._idx_sorted = _empty(self.n, torch.int32, self.device)
            L.call("vapc_record_indices", _ptr(self.xyzw), _ptr(self.pos_sorted), self.n, _ptr(self._idx_sorted), _stream())
        return self._idx_sorted

    def covariance_reference_formula(self) -> torch.Tensor:
        """[6, V] covariance (xx xy xz yy yz zz) exactly as the reference's compute_covariance_matrix produces it (vapc.py:947-1006):
        raw-moment formula over pandas' Kahan group sums, bit for bit — including its loss of digits at large coordinates.  For
        diffing against the reference; `stats(cov=True)` is the accurate one."""
        V = self.nvoxels
        out = torch.empty((6, V), dtype=torch.float64, device=self.device)
        L.call("vapc_cov_reference_formula", _ptr(self.xyzw), _ptr(self.pos_sorted), _ptr(self.start), V, _ptr(out), _stream())
        return out

    def centroid_reference_formula(self) -> torch.Tensor:
        """[3, V] centroid exactly as the reference's compute_center_of_gravity produces it (vapc.py:843-850): pandas' groupby mean =
        the Kahan sum of the voxel's coordinates in row order / count, bit for bit.  Distances to it, their per-voxel minimum and
        the rows tied at that minimum are then the reference's own (vapc.py:790-825, :1195)."""
        V = self.nvoxels
        out = torch.empty((3, V), dtype=torch.float64, device=self.device)
        L.call("vapc_cog_reference_formula", _ptr(self.xyzw), _ptr(self.pos_sorted), _ptr(self.start), V, _ptr(out), _stream())
        return out

    def std_reference_formula(self) -> torch.Tensor:
        """[3, V] standard deviation exactly as the reference's compute_std_of_cog produces it (vapc.py:873-880): pandas' groupby std =
        Welford's recurrence in row order, / (n - 1), square root — bit for bit, its loss of digits at large coordinates included.
        For diffing against the reference; `stats(std=True)` is the accurate one."""
        V = self.nvoxels
        out = torch.empty((3, V), dtype=torch.float64, device=self.device)
        L.call("vapc_std_reference_formula", _ptr(self.xyzw), _ptr(self.pos_sorted), _ptr(self.start), V, _ptr(out), _stream())
        return out

    @property
    def centroid_formula(self) -> str:
        """Which double `stats(cog=True).cog` — and with it point_distance() / closest_to_cog() — carries: "exact" (the mean from the
        shifted moments of the reduction, default) or "reference" (centroid_reference_formula: the drop-in facade's choice)."""
        return self._centroid_formula

    @centroid_formula.setter
    def centroid_formula(self, value: str):
        if value not in ("exact", "reference"):
            raise ValueError("centroid_formula must be 'exact' or 'reference'")
        if value != self._centroid_formula:
            self._stats_cache.pop("cog", None)
        self._centroid_formula = value

    # ------------------------------------------------------------------------------------
    def stats(self, density=False, cog=False, std=False, cov=False, eig=False, eig_n=False, feat=False) -> VoxelStats:
        """One launch of vapc_voxel_stats for whatever is not cached yet."""
        V, dev = self.nvoxels, self.device
        want = dict(density=density, cog=cog, std=std, cov=cov, eig=eig, eig_n=eig_n, feat=feat)
        shapes = dict(density=(V,), cog=(3, V), std=(3, V), cov=(6, V), eig=(3, V), eig_n=(3, V), feat=(8, V))
        new = {k: torch.empty(shapes[k], dtype=torch.float64, device=dev) for k, w in want.items() if w and k not in self._stats_cache}
        if new:
            # the kernel forms everything from the (count, mean, M2) moments; a "reference" centroid comes from its own pass instead
            outs = [new.get(k) if k != "cog" or self._centroid_formula == "exact" else None for k in ("density", "cog", "std", "cov", "eig", "eig_n", "feat")]
            if any(o is not None for o in outs):
                order = (C.c_int32 * 3)(*self.axis_order)
                L.call("vapc_voxel_stats_axes", _ptr(self.voxel_keys), _ptr(self.count), _ptr(self.mean3), _ptr(self.m2), V, C.byref(self.grid), order,
                       *[_ptr(o) for o in outs], _stream())
            if "cog" in new and self._centroid_formula == "reference":
                new["cog"] = self.centroid_reference_formula()
            self._stats_cache.update(new)
        return VoxelStats(**{k: self._stats_cache[k] for k, w in want.items() if w})

    def voxel_coords(self):
        """(vx, vy, vz) int64 [V] of the voxel table rows."""
        V, dev = self.nvoxels, self.device
        out = [_empty(V, torch.int64, dev) for _ in range(3)]
        L.call("vapc_unpack_keys", _ptr(self.voxel_keys), V, C.byref(self.grid), _ptr(out[0]), _ptr(out[1]), _ptr(out[2]), _stream())
        return out

    def center_corner(self, center=True, corner=True):
        V, dev = self.nvoxels, self.device
        c = torch.empty((3, V), dtype=torch.float64, device=dev) if center else None
        k = torch.empty((3, V), dtype=torch.float64, device=dev) if corner else None
        L.call("vapc_voxel_center_corner", _ptr(self.voxel_keys), V, C.byref(self.grid), _ptr(c), _ptr(k), _stream())
        return c, k

    def closest_to_cog(self, cog: Optional[torch.Tensor] = None):
        """(min_distance [V] f64, argmin original point index [V] int32-as-uint32).  cog ([3, V], optional): centroids to
        measure against instead of this cloud's own (multi-GPU: the merged table's centroids of the rank's partial rows)."""
        V, dev = self.nvoxels, self.device
        cog = self.stats(cog=True).cog if cog is None else cog.contiguous()
        md = _empty(V, torch.float64, dev)
        am = _empty(V, torch.int32, dev)
        L.call("vapc_closest_to_cog", _ptr(self.keys_sorted), _ptr(self.pos_sorted), self.n, self.grid.key_bytes, _ptr(self.xyzw), _ptr(cog), V,
               _ptr(md), _ptr(am), _ptr(self._seg_ws), self._seg_ws.numel(), _stream())
        return md, am

    def point_distance(self):
        """distance of every point to its voxel's centroid, original point order (vapc.py:790-794)."""
        cog = self.stats(cog=True).cog
        out = _empty(self.n, torch.float64, self.device)
        L.call("vapc_point_distance", _ptr(self.x), _ptr(self.y), _ptr(self.z), self.n, _ptr(self.point2voxel), _ptr(cog), self.nvoxels,
               _ptr(out), _stream())
        return out

    def broadcast(self, table: torch.Tensor) -> torch.Tensor:
        """table[point2voxel] for a [V] fp64 / int32 voxel column (the reference's per-point broadcast)."""
        out = _empty(self.n, table.dtype, self.device)
        if table.dtype == torch.float64:
            L.call("vapc_gather_f64", _ptr(table), _ptr(self.point2voxel), self.n, _ptr(out), _stream())
        elif table.dtype == torch.int32:
            L.call("vapc_gather_u32", _ptr(table), _ptr(self.point2voxel), self.n, _ptr(out), _stream())
        else:
            raise TypeError(table.dtype)
        return out

    def mode_count(self, attr, thresholds: Sequence[float] = (), want_mode=True):
        """mode and mode_count,p of an integer-valued attribute column (vapc.py:407-433).
        Returns (mode [V] int64 or None, {p: mode_count [V] int64})."""
        dev = self.device
        a = _dev_f64(attr, dev)
        if a.numel() != self.n:
            raise ValueError("attribute length differs from the cloud")
        # values np.bincount(x.astype(int)) rejects — negative, NaN, infinite (vapc.py:409, :423): `mode_count` raises for such a
        # voxel, `mode` falls back to the histogram mode of that voxel alone (vapc.py:411-412); the other voxels are not affected
        rejected = ~(a >= 0) | torch.isinf(a)
        flagged_rows = None
        if bool(rejected.any()):
            if list(thresholds):
                raise ValueError("Cannot compute 'mode_count' for negative or non-finite attribute values")
            flagged_rows = torch.unique(_u32_t(self.point2voxel)[rejected])
            a_all, a = a, torch.where(rejected, torch.zeros_like(a), a)
        amax = float(a.max().item()) if self.n else 0.0
        attr_bits = max(1, int(max(amax, 0.0)).bit_length())
        if attr_bits > 32:
            raise ValueError("attribute values too large for mode / mode_count")
        vbits = max(1, int(self.nvoxels - 1).bit_length())
        # large voxels (hundreds of points each): the one-thread-per-voxel run scan of vapc_mode_count would walk them serially;
        # the (voxel, value) runs are encoded by a point-parallel pass instead and evaluated per voxel (vapc_mode_count_runs)
        by_runs = self.n > 256 * self.nvoxels
        kb = 4 if vbits + attr_bits <= 32 and not by_runs else 8  # (voxel row, value) keys of <= 32 bits: half the sort traffic
        keys = _empty(self.n, torch.int32 if kb == 4 else torch.int64, dev)
        status = torch.zeros(1, dtype=torch.int32, device=dev)
        L.call("vapc_mode_keys", _ptr(self.point2voxel), _ptr(a), self.n, attr_bits, kb, _ptr(keys), _ptr(status), _stream())
        ks, _ = sort_pairs(keys, vbits + attr_bits)
        if int(status.item()) & 2:
            raise ValueError("Cannot compute 'mode' / 'mode_count' for negative or non-finite attribute values")
        mode = _empty(self.nvoxels, torch.int64, dev) if want_mode else None
        counts = {}
        todo = list(thresholds) or [None]
        if by_runs:
            run_keys, run_weight = run_lengths(ks)
        for i, p in enumerate(todo):
            mc = _empty(self.nvoxels, torch.int64, dev) if p is not None else None
            if by_runs:
                L.call("vapc_mode_count_runs", _ptr(run_keys), _ptr(run_weight), run_keys.numel(), attr_bits, self.nvoxels,
                       float(p if p is not None else 0.0), _ptr(mode if i == 0 else None), _ptr(mc), _stream())
                if p is not None:
                    counts[p] = mc
                continue
            L.call("vapc_mode_count", _ptr(ks), self.n, attr_bits, kb, _ptr(self.start), self.nvoxels, float(p if p is not None else 0.0),
                   _ptr(mode if i == 0 else None), _ptr(mc), _stream())
            if p is not None:
                counts[p] = mc
        if flagged_rows is not None and mode is not None:
            # the reference's per-voxel fallback for the flagged voxels only: their points come out of the sorted order in one
            # gather, all histogram modes in one batch of tensor ops (segmented_histogram_modes)
            mode = mode.to(torch.float64)  # a column holding histogram modes is a float column
            start = _u32_t(self.start)
            lens = start[flagged_rows + 1] - start[flagged_rows]
            seg = torch.repeat_interleave(torch.arange(flagged_rows.numel(), device=dev), lens)
            within = torch.arange(int(lens.sum().item()), device=dev) - (torch.cumsum(lens, 0) - lens)[seg]
            vals = a_all[_u32_t(self.idx_sorted)[start[flagged_rows][seg] + within]]
            modes = segmented_histogram_modes(vals, lens)
            if modes is None:
                # a voxel the reference's histogram raises for (one value, zero inter-quartile range, NaN), or bins beyond the
                # batch's memory bound: voxel by voxel on the host, in row order like the groupby, raising what numpy raises
                from .utilities import compute_mode_continuous

                host = vals.cpu().numpy()
                cuts = np.concatenate([[0], np.cumsum(lens.cpu().numpy())])
                modes = torch.tensor([compute_mode_continuous(host[cuts[i]:cuts[i + 1]]) for i in range(len(cuts) - 1)],
                                     dtype=torch.float64, device=dev)
            mode[flagged_rows] = modes
        return mode, counts

    def percentage_occupied(self) -> float:
        """round(V / bbox_cells * 100, 2)  (vapc.py:761-773)."""
        vx, vy, vz = self.voxel_coords()
        ext = [int(v.max().item()) - int(v.min().item()) + 1 for v in (vx, vy, vz)]
        return round(self.nvoxels / (ext[0] * ext[1] * ext[2]) * 100, 2)


def segmented_histogram_modes(values: torch.Tensor, lengths: torch.Tensor, max_bins: int = 1 << 27) -> Optional[torch.Tensor]:
    """compute_mode_continuous (the reference's utilities.py:114-159) of many samples at once: `values` holds the samples back to
    back, `lengths` [S] their sizes; returns the S modes (fp64), or None when the batch cannot take this route (a non-finite value,
    a degenerate bin width — the reference raises there, and the caller's per-sample route raises the same error —, or more than
    `max_bins` bins in total).  Every step restates numpy's own arithmetic operation by operation in IEEE fp64 tensor ops, so the
    result equals np.percentile / np.histogram / np.linspace bit for bit: the samples sorted inside their segments (two stable sorts),
    the quartiles by numpy's `linear` rule (virtual index (n - 1) q, _lerp with its t >= 0.5 branch), bins = ceil(range / (2 iqr /
    n ** (1/3))) — n ** (1/3) through Python's pow on the host, once per distinct n: the device's pow may differ in the last bit —,
    edges = arange * (range / bins) + min with the last edge set to max, numpy's bin index with its one-ULP corrections, density =
    counts / diff(edges) / n, first maximum, centre of that bin."""
    dev = values.device
    S = lengths.numel()
    if S == 0:
        return torch.empty(0, dtype=torch.float64, device=dev)
    if not bool(torch.isfinite(values).all()):
        return None
    lengths = lengths.to(torch.int64)
    off = torch.cumsum(lengths, 0) - lengths
    seg = torch.repeat_interleave(torch.arange(S, device=dev), lengths)
    # samples in ascending order inside their segments
    o1 = torch.sort(values, stable=True).indices
    o2 = torch.sort(seg[o1], stable=True).indices
    sv = values[o1[o2]]
    last = off + lengths - 1
    vmin, vmax = sv[off], sv[last]
    nf = lengths.to(torch.float64)

    def quartile(q):
        virt = (nf - 1.0) * q
        prev = torch.floor(virt)
        above = virt >= nf - 1.0
        pi = torch.where(above, lengths - 1, prev.to(torch.int64))
        ni = torch.where(above, lengths - 1, prev.to(torch.int64) + 1)
        t = virt - torch.where(above, torch.full_like(prev, -1.0), prev)
        a, b = sv[off + pi], sv[off + ni]
        d = b - a
        return torch.where(t >= 0.5, b - d * (1.0 - t), a + d * t)

    iqr = quartile(0.75) - quartile(0.25)
    uniq, inv = torch.unique(lengths, return_inverse=True)
    cbrt = torch.tensor([int(c) ** (1 / 3) for c in uniq.cpu().tolist()], dtype=torch.float64).to(dev)[inv]
    ratio = (vmax - vmin) / (2.0 * iqr / cbrt)
    if not bool(torch.isfinite(ratio).all()):
        return None
    bins = torch.ceil(ratio).to(torch.int64)
    if bool((bins < 1).any()) or int(bins.sum().item()) + S > max_bins:
        return None
    # bin edges of every segment back to back (bins + 1 each): np.linspace(min, max, bins + 1)
    ne = bins + 1
    eoff = torch.cumsum(ne, 0) - ne
    eseg = torch.repeat_interleave(torch.arange(S, device=dev), ne)
    j = (torch.arange(int(ne.sum().item()), device=dev) - eoff[eseg]).to(torch.float64)
    delta = vmax - vmin
    binsf = bins.to(torch.float64)
    step = delta / binsf
    edges = torch.where(step[eseg] == 0, j / binsf[eseg] * delta[eseg], j * step[eseg]) + vmin[eseg]
    edges[eoff + bins] = vmax
    # numpy's index of every sample: trunc((v - first) / (last - first) * bins), then the corrections around the edges
    segs = seg  # the sorted samples keep the segment layout
    f = (sv - vmin[segs]) / delta[segs] * binsf[segs]
    idx = f.to(torch.int64)
    nb = bins[segs]
    idx = torch.where(idx == nb, idx - 1, idx)
    eo = eoff[segs]
    idx = torch.where(sv < edges[eo + idx], idx - 1, idx)
    idx = torch.where((sv >= edges[eo + idx + 1]) & (idx != nb - 1), idx + 1, idx)
    boff = eoff - torch.arange(S, device=dev)  # first bin of each segment among all bins
    nbins_total = int(bins.sum().item())
    counts = torch.bincount(boff[segs] + idx, minlength=nbins_total)
    bseg = torch.repeat_interleave(torch.arange(S, device=dev), bins)
    bj = torch.arange(nbins_total, device=dev) - boff[bseg]
    lo, hi = edges[eoff[bseg] + bj], edges[eoff[bseg] + bj + 1]
    dens = counts.to(torch.float64) / (hi - lo) / nf[bseg]
    best = torch.full((S,), float("-inf"), dtype=torch.float64, device=dev).scatter_reduce(0, bseg, dens, "amax", include_self=True)
    first = torch.full((S,), nbins_total, dtype=torch.int64, device=dev).scatter_reduce(
        0, bseg, torch.where(dens == best[bseg], bj, torch.full_like(bj, nbins_total)), "amin", include_self=True)
    k = eoff + first
    return (edges[k] + edges[k + 1]) / 2.0


def run_lengths(keys_sorted: torch.Tensor, weights: Optional[torch.Tensor] = None):
    """(distinct keys, summed weights) of sorted 64-bit keys (vapc_run_lengths); weights: int32 per element or None = 1."""
    dev = keys_sorted.device
    n = keys_sorted.numel()
    if n == 0:
        return _empty(0, torch.int64, dev), _empty(0, torch.int64, dev)
    ws_bytes = L.raw("vapc_segment_workspace_bytes")(n)
    ws = _empty(ws_bytes, torch.uint8, dev)
    nv = torch.zeros(1, dtype=torch.int64, device=dev)
    L.call("vapc_count_voxels", _ptr(keys_sorted), n, 8, _ptr(nv), _ptr(ws), ws_bytes, _stream())
    r = int(nv.item())
    run_keys = _empty(r, torch.int64, dev)
    run_weight = _empty(r, torch.int64, dev)
    L.call("vapc_run_lengths", _ptr(keys_sorted), _ptr(weights), n, r, _ptr(run_keys), _ptr(run_weight), _ptr(ws), ws_bytes, _stream())
    return run_keys, run_weight


def attribute_statistics(cloud: VoxelCloud, attr, median=True) -> Dict[str, torch.Tensor]:
    """mean / sum / min / max / var / std (ddof = 0) / median of an fp64 attribute per voxel
    (vapc.py:393-406); [V] device tensors."""
    dev, V, n = cloud.device, cloud.nvoxels, cloud.n
    a = _dev_f64(attr, dev)
    if a.numel() != n:
        raise ValueError("attribute length differs from the cloud")
    out = {k: _empty(V, torch.float64, dev) for k in ("sum", "mean", "min", "max", "var")}
    L.call("vapc_attr_stats", _ptr(a), _ptr(cloud.idx_sorted), _ptr(cloud.start), V, _ptr(out["sum"]), _ptr(out["mean"]), _ptr(out["min"]),
           _ptr(out["max"]), _ptr(out["var"]), _stream())
    out["std"] = torch.sqrt(out["var"])
    if median:
        keys = _empty(n, torch.int64, dev)
        L.call("vapc_f64_sort_keys", _ptr(a), n, _ptr(keys), _stream())
        _, by_value = sort_pairs(keys, 64)
        rows = _empty(n, torch.int32, dev)
        L.call("vapc_gather_u32", _ptr(cloud.point2voxel), _ptr(by_value), n, _ptr(rows), _stream())
        _, by_voxel_value = sort_pairs(rows, max(1, int(V - 1).bit_length()), idx=by_value)
        out["median"] = _empty(V, torch.float64, dev)
        L.call("vapc_attr_median", _ptr(a), _ptr(by_voxel_value), _ptr(cloud.start), V, _ptr(out["median"]), _stream())
    return out


def lexsort_unique_rows(x, y, z, device=None, as_tensor=False):
    """Row numbers that `df.drop_duplicates(subset=[X, Y, Z])` followed by `df.groupby([X, Y, Z]).<agg>()` visits, in the
    order of the result (vapc.py:1208-1209): rows sorted lexicographically by (X, Y, Z), of rows with identical coordinates only
    the first (lowest row number), rows with a NaN coordinate dropped.  Three stable 64-bit radix sorts on the GPU (Z, then
    Y, then X) instead of pandas' hash + sort on the host.  as_tensor: the rows as an int64 device tensor instead of a host array."""
    _require_cuda()
    if device is None:
        device = torch.device("cuda", torch.cuda.current_device())
    cols = [_dev_f64(a, device) + 0.0 for a in (x, y, z)]  # + 0.0: -0.0 and 0.0 are one group, as in pandas
    n = cols[0].numel()
    if n == 0:
        return torch.zeros(0, dtype=torch.int64, device=device) if as_tensor else np.zeros(0, dtype=np.int64)
    order = None
    for c in reversed(cols):
        keys = _empty(n, torch.int64, device)
        src = c if order is None else c[order.to(torch.int64) & 0xFFFFFFFF]
        L.call("vapc_f64_sort_keys", _ptr(src.contiguous()), n, _ptr(keys), _stream())
        _, order = sort_pairs(keys, 64, idx=order)
    rows = order.to(torch.int64) & 0xFFFFFFFF
    sx, sy, sz = (c[rows] for c in cols)
    keep = torch.ones(n, dtype=torch.bool, device=device)
    keep[1:] = (sx[1:] != sx[:-1]) | (sy[1:] != sy[:-1]) | (sz[1:] != sz[:-1])
    keep &= ~(torch.isnan(sx) | torch.isnan(sy) | torch.isnan(sz))
    return rows[keep] if as_tensor else rows[keep].cpu().numpy()


# ----------------------------------------------------------------------------------------
# masks
# ----------------------------------------------------------------------------------------
class VoxelMask:
    """Hash set of voxel triples on the device (MultiIndex.isin, vapc.py:594-598)."""

    def __init__(self, mvx, mvy, mvz, voxel_size, origin, device=None):
        _require_cuda()
        if device is None:
            device = torch.device("cuda", torch.cuda.current_device())
        mv = [torch.as_tensor(np.asarray(a, dtype=np.int64) if not isinstance(a, torch.Tensor) else a).to(device=device, dtype=torch.int64).contiguous()
              for a in (mvx, mvy, mvz)]
        self.n = mv[0].numel()
        self.device = device
        if self.n == 0:
            self.grid = grid_from_voxel_bounds(voxel_size, origin, (0, 0, 0), (0, 0, 0))
        else:
            lo = [int(a.min().item()) for a in mv]
            hi = [int(a.max().item()) for a in mv]
            self.grid = grid_from_voxel_bounds(voxel_size, origin, lo, hi)
        if self.grid.total_bits > 63:
            raise L.VapcError(L.VAPC_E_KEYSPACE, "mask extent needs more than 63 key bits")
        keys = _empty(self.n, torch.int64, device)
        status = torch.zeros(1, dtype=torch.int32, device=device)
        L.call("vapc_pack_triples_i64", _ptr(mv[0]), _ptr(mv[1]), _ptr(mv[2]), self.n, C.byref(self.grid), _ptr(keys), _ptr(status), _stream())
        cap = C.c_int64(0)
        nbytes = L.raw("vapc_mask_table_bytes")(self.n, C.byref(cap))
        self.capacity = cap.value
        self.table = _empty(nbytes, torch.uint8, device)
        L.call("vapc_mask_build", _ptr(keys), self.n, _ptr(self.table), self.capacity, _stream())

    def contains_points(self, x, y, z) -> torch.Tensor:
        """uint8 [N]: 1 where the point's voxel (same voxel_size / origin as the mask) is in the set."""
        x, y, z = (_dev_f64(a, self.device) for a in (x, y, z))
        n = x.numel()
        out = _empty(n, torch.uint8, self.device)
        L.call("vapc_mask_lookup_points", _ptr(x), _ptr(y), _ptr(z), n, C.byref(self.grid), _ptr(self.table), self.capacity, _ptr(out), _stream())
        return out

    def contains_voxels(self, vx, vy, vz) -> torch.Tensor:
        v = [torch.as_tensor(np.asarray(a, dtype=np.int64) if not isinstance(a, torch.Tensor) else a).to(device=self.device, dtype=torch.int64).contiguous()
             for a in (vx, vy, vz)]
        n = v[0].numel()
        keys = _empty(n, torch.int64, self.device)
        status = torch.zeros(1, dtype=torch.int32, device=self.device)
        L.call("vapc_pack_triples_i64", _ptr(v[0]), _ptr(v[1]), _ptr(v[2]), n, C.byref(self.grid), _ptr(keys), _ptr(status), _stream())
        out = _empty(n, torch.uint8, self.device)
        L.call("vapc_mask_lookup_keys", _ptr(keys), n, _ptr(self.table), self.capacity, _ptr(out), _stream())
        return out


def unique_voxels_first_occurrence(vx, vy, vz, device=None):
    """drop_duplicates() on voxel triples: unique rows in first-occurrence order (vapc.py:507-510,
    :524-525).  Packs, stable-sorts (key, position), takes each run's first position, then sorts those
    positions.  Returns three int64 device columns."""
    _require_cuda()
    if device is None:
        device = torch.device("cuda", torch.cuda.current_device())
    v = [torch.as_tensor(np.asarray(a, dtype=np.int64) if not isinstance(a, torch.Tensor) else a).to(device=device, dtype=torch.int64).contiguous()
         for a in (vx, vy, vz)]
    n = v[0].numel()
    if n == 0:
        return v
    lo = [int(a.min().item()) for a in v]
    hi = [int(a.max().item()) for a in v]
    g = grid_from_voxel_bounds(1.0, (0.0, 0.0, 0.0), lo, hi)
    if g.total_bits > 63:
        raise L.VapcError(L.VAPC_E_KEYSPACE, "voxel extent needs more than 63 key bits")
    keys = _empty(n, torch.int64, device)
    status = torch.zeros(1, dtype=torch.int32, device=device)
    L.call("vapc_pack_triples_i64", _ptr(v[0]), _ptr(v[1]), _ptr(v[2]), n, C.byref(g), _ptr(keys), _ptr(status), _stream())
    ks, order = sort_pairs(keys, g.total_bits)
    ws_bytes = L.raw("vapc_segment_workspace_bytes")(n)
    ws = _empty(ws_bytes, torch.uint8, device)
    nv = torch.zeros(1, dtype=torch.int64, device=device)
    L.call("vapc_count_voxels", _ptr(ks), n, 8, _ptr(nv), _ptr(ws), ws_bytes, _stream())
    # heads of runs: stable sort => the first element of each run is the first occurrence
    head = torch.ones(n, dtype=torch.bool, device=device)
    head[1:] = ks[1:] != ks[:-1]
    first_pos = order[head]  # uint32 positions stored as int32
    assert first_pos.numel() == int(nv.item())
    pos_sorted, _ = sort_pairs(first_pos.clone(), max(1, int(n - 1).bit_length()))
    sel = pos_sorted.to(torch.int64) & 0xFFFFFFFF
    return [a[sel] for a in v]


def buffer_expand(vx, vy, vz, buffer_size: int, device=None):
    """Every triple + all (2b+1)^3 offsets in the reference's order (vapc.py:513-521); [3, n*(2b+1)^3]."""
    _require_cuda()
    if device is None:
        device = torch.device("cuda", torch.cuda.current_device())
    v = [torch.as_tensor(np.asarray(a, dtype=np.int64) if not isinstance(a, torch.Tensor) else a).to(device=device, dtype=torch.int64).contiguous()
         for a in (vx, vy, vz)]
    n = v[0].numel()
    w = 2 * int(buffer_size) + 1
    out = torch.empty((3, n * w ** 3), dtype=torch.int64, device=device)
    L.call("vapc_buffer_expand", _ptr(v[0]), _ptr(v[1]), _ptr(v[2]), n, int(buffer_size), _ptr(out), _stream())
    return out


# ----------------------------------------------------------------------------------------
# two-epoch change detection
# ----------------------------------------------------------------------------------------
def join_voxel_tables(v1: Sequence[torch.Tensor], v2: Sequence[torch.Tensor]):
    """Inner join + the two anti-joins of two voxel tables given as (vx, vy, vz) int64 device columns
    whose rows are UNIQUE (bi_vapc.py:84, :30-31).  Returns (i1, i2, only1, only2) int64 index tensors:
    pairs in epoch-1 row order; only1 / only2 in lexicographic voxel order (Index.difference sorts)."""
    dev = v1[0].device
    n1, n2 = v1[0].numel(), v2[0].numel()
    lo = [min(int(a.min().item()) if n1 else 0, int(b.min().item()) if n2 else 0) for a, b in zip(v1, v2)]
    hi = [max(int(a.max().item()) if n1 else 0, int(b.max().item()) if n2 else 0) for a, b in zip(v1, v2)]
    g = grid_from_voxel_bounds(1.0, (0.0, 0.0, 0.0), lo, hi)
    if g.total_bits > 63:
        raise L.VapcError(L.VAPC_E_KEYSPACE, "joined voxel extent needs more than 63 key bits")
    sorted_keys, orders = [], []
    for v, n in ((v1, n1), (v2, n2)):
        keys = _empty(n, torch.int64, dev)
        status = torch.zeros(1, dtype=torch.int32, device=dev)
        L.call("vapc_pack_triples_i64", _ptr(v[0]), _ptr(v[1]), _ptr(v[2]), n, C.byref(g), _ptr(keys), _ptr(status), _stream())
        ks, order = sort_pairs(keys, g.total_bits)
        sorted_keys.append(ks)
        orders.append(order.to(torch.int64) & 0xFFFFFFFF)
    m1 = _empty(n1, torch.int32, dev)
    m2 = _empty(n2, torch.int32, dev)
    L.call("vapc_join_sorted_keys", _ptr(sorted_keys[0]), n1, _ptr(sorted_keys[1]), n2, _ptr(m1), _ptr(m2), _stream())
    m1 = m1.to(torch.int64) & 0xFFFFFFFF
    m2 = m2.to(torch.int64) & 0xFFFFFFFF
    miss = 0xFFFFFFFF
    # back to original row numbering: sorted position p of table 1 is row orders[0][p]
    match_row1 = torch.full((n1,), -1, dtype=torch.int64, device=dev)
    hit = m1 != miss
    match_row1[orders[0][hit]] = orders[1][m1[hit]]
    i1 = torch.nonzero(match_row1 >= 0).flatten()
    i2 = match_row1[i1]
    only1 = orders[0][m1 == miss]  # already in key (lexicographic) order
    only2 = orders[1][m2 == miss]
    return i1, i2, only1, only2


def mahalanobis(cog1, cov1, cog2, cov2, i1, i2, alpha=0.005):
    """bi_vapc.py:86-138 on joined rows.  cog*: [3, V] f64, cov*: [6, V] f64, i1 / i2: int64 rows.
    Returns dict of device tensors: distance, d1, d2, p_value, significance, change_type."""
    dev = cog1.device
    r = i1.numel()
    a = i1.to(torch.int32).contiguous()
    b = i2.to(torch.int32).contiguous()
    out = dict(distance=_empty(r, torch.float64, dev), d1=_empty(r, torch.float64, dev), d2=_empty(r, torch.float64, dev),
               p_value=_empty(r, torch.float64, dev), significance=_empty(r, torch.int32, dev), change_type=_empty(r, torch.int32, dev))
    L.call("vapc_mahalanobis", _ptr(cog1.contiguous()), _ptr(cov1.contiguous()), cog1.shape[1], _ptr(cog2.contiguous()), _ptr(cov2.contiguous()),
           cog2.shape[1], _ptr(a), _ptr(b), r, float(alpha), _ptr(out["distance"]), _ptr(out["d1"]), _ptr(out["d2"]), _ptr(out["p_value"]),
           _ptr(out["significance"]), _ptr(out["change_type"]), _stream())
    return out


# ----------------------------------------------------------------------------------------
# neighbour queries on a voxelised point set
# ----------------------------------------------------------------------------------------
_MAX_AXIS_CELLS = float(1 << 20)  # cell size floor: three axes of <= 20 bits keep the packed key within 63 bits


def _cell_cloud(x, y, z, cell_size, origin=(0.0, 0.0, 0.0)) -> VoxelCloud:
    """The searched set voxelised with `cell_size`: its voxel table is the cell list of vapc_cell_nn /
    vapc_cluster_components.  The cell size is raised when the extent would need more than 2^20 cells per axis."""
    _require_cuda()
    dev = x.device if isinstance(x, torch.Tensor) and x.is_cuda else torch.device("cuda", torch.cuda.current_device())
    x, y, z = (_dev_f64(a, dev) for a in (x, y, z))
    bounds = coordinate_bounds(x, y, z)
    extent = float(max(bounds[3] - bounds[0], bounds[4] - bounds[1], bounds[5] - bounds[2]))
    h = max(float(cell_size), extent / _MAX_AXIS_CELLS * 1.0000001)
    return VoxelCloud.build(x, y, z, h, origin, bounds=bounds, want_point2voxel=False, keep_columns=False)


def nearest_neighbors(qx, qy, qz, cells: VoxelCloud):
    """Exact nearest point of the voxelised set `cells` for every query point (KDTree(set).query(q, k=1),
    bi_vapc.py:179-184).  Returns (index int64 [nq], distance f64 [nq]) device tensors."""
    dev = cells.device
    qx, qy, qz = (_dev_f64(a, dev) for a in (qx, qy, qz))
    nq = qx.numel()
    idx = _empty(nq, torch.int32, dev)
    dist = _empty(nq, torch.float64, dev)
    L.call("vapc_cell_nn", _ptr(qx), _ptr(qy), _ptr(qz), nq, C.byref(cells.grid), _ptr(cells.voxel_keys), cells.nvoxels, _ptr(cells.start),
           _ptr(cells.pos_sorted), _ptr(cells.xyzw), cells.n, 1, _ptr(idx), _ptr(dist), _stream())
    # queries whose neighbour is many cells away (sets far apart, isolated outliers): a grid search degenerates there, a tiled
    # all-pairs pass over just those queries has a bounded cost
    far = torch.nonzero(idx == -1).flatten()
    if far.numel():
        L.call("vapc_nn_all_pairs", _ptr(qx), _ptr(qy), _ptr(qz), _ptr(far), far.numel(), _ptr(cells.xyzw), cells.n, _ptr(idx), _ptr(dist), _stream())
    return idx.to(torch.int64) & 0xFFFFFFFF, dist


def mutual_nearest_neighbors(c1: Sequence, c2: Sequence, cell_sizes=(None, None), origins=((0.0, 0.0, 0.0), (0.0, 0.0, 0.0))):
    """bi_vapc.py:173-186 + :62-64: rows of set 1 whose nearest row of set 2 has them as ITS nearest row of set 1.
    c1 / c2: (x, y, z) columns.  cell_sizes: cell size of the search grid over set 1 / set 2 (None: about two points
    per cell of the bounding box).  Returns (idx1, idx2, distance) device tensors in set-1 order."""
    _require_cuda()
    dev = torch.device("cuda", torch.cuda.current_device())
    sets = [[_dev_f64(a, dev) for a in c] for c in (c1, c2)]
    clouds = []
    for s, h, o in zip(sets, cell_sizes, origins):
        if h is None:
            b = coordinate_bounds(*s)
            ext = np.maximum(np.asarray(b[3:]) - np.asarray(b[:3]), 1e-12)
            h = float(np.cbrt(2.0 * np.prod(ext) / s[0].numel()))
        clouds.append(_cell_cloud(*s, h, o))
    fwd, dist = nearest_neighbors(*sets[0], clouds[1])
    back, _ = nearest_neighbors(*sets[1], clouds[0])
    n1 = sets[0][0].numel()
    idx1 = torch.nonzero(back[fwd] == torch.arange(n1, device=dev)).flatten()
    return idx1, fwd[idx1], dist[idx1]


def cluster_components(cx, cy, cz, cluster_distance: float, integer_rows: bool = False):
    """compute_clusters of the reference (vapc.py:1324-1357) as connected components of the graph "row distance <=
    cluster_distance", with the reference's label numbering: in its sequential sweep a row opens a new label iff no
    earlier row lies within two hops (vapc_cluster_components' hop2_first), labels merge to the lowest, so a component
    carries the label opened by its lowest row.  Identical to the reference whenever no row has more than 50 rows
    (its `nr_of_neighbours`, itself included) within cluster_distance; beyond that the reference truncates its
    neighbourhoods to a KD-tree dependent subset and may split what is connected here.
    integer_rows: the columns hold integers (voxel indices): identical rows are collapsed into one unit standing for
    all its points first.  Returns (cluster_id int64 [n], cluster_size int64 [n], openings) on the device."""
    _require_cuda()
    dev = torch.device("cuda", torch.cuda.current_device())
    cols = [_dev_f64(a, dev) for a in (cx, cy, cz)]
    d = float(cluster_distance)
    first = weight = p2v = None
    if integer_rows:
        units = VoxelCloud.build(*cols, 1.0, (0.0, 0.0, 0.0), keep_columns=False)
        cols = [v.to(torch.float64) for v in units.voxel_coords()]
        first, weight, p2v = units.first_idx, units.count, units.point2voxel.to(torch.int64) & 0xFFFFFFFF
        del units
    n = cols[0].numel()
    cells = _cell_cloud(*cols, d * (1.0 + 2.0 ** -30) if d > 0 else 1.0)
    # every unit visits the units of its 27 cells: ~27 n^2 / cells distance tests.  A cluster distance that spans most of the cloud
    # makes that quadratic (the reference's sweep is slower still); refuse instead of occupying the GPU for hours
    if 27.0 * n * n / max(cells.nvoxels, 1) > 2e13:
        raise ValueError(f"compute_clusters: cluster_distance {d} spans too much of the cloud ({n} rows in {cells.nvoxels} cells of that size): "
                         "the neighbourhood search would be quadratic")
    root, hop2, comp_first = (_empty(n, torch.int32, dev) for _ in range(3))
    comp_size = _empty(n, torch.int64, dev)
    ws_bytes = L.raw("vapc_cluster_workspace_bytes")(n)
    ws = _empty(ws_bytes, torch.uint8, dev)
    L.call("vapc_cluster_components", C.byref(cells.grid), _ptr(cells.voxel_keys), cells.nvoxels, _ptr(cells.start), _ptr(cells.pos_sorted),
           _ptr(cells.xyzw), n, d, _ptr(first), _ptr(weight), _ptr(root), _ptr(hop2), _ptr(comp_first), _ptr(comp_size), _ptr(ws), ws_bytes, _stream())
    u32 = lambda t: t.to(torch.int64) & 0xFFFFFFFF  # noqa: E731
    first64 = u32(first) if first is not None else torch.arange(n, device=dev)
    root = u32(root)
    opens = torch.sort(first64[u32(hop2) == first64]).values
    label = torch.searchsorted(opens, u32(comp_first)[root]) + 1
    size = comp_size[root]
    if p2v is not None:
        label, size = label[p2v], size[p2v]
    return label, size, int(opens.numel())
