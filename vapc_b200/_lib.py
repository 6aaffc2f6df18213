"""ctypes binding of libvapc_b200.so (the C ABI declared in include/vapc_b200.h).

The library is built in-tree by `vapc_b200/csrc/build.sh` (or `__graft_entry__.build()`).
There is no CPU fallback: if the shared library is missing, importing this module raises.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("VAPC_B200_LIB") or os.path.join(_HERE, "libvapc_b200.so")  # env: a development build

VAPC_OK = 0
VAPC_E_INVALID = -1
VAPC_E_CUDA = -2
VAPC_E_KEYSPACE = -3
VAPC_E_WORKSPACE = -4
VAPC_E_RANGE = -5


class VapcError(RuntimeError):
    """A libvapc_b200 call failed (CUDA error, workspace, key space)."""

    def __init__(self, code, msg):
        super().__init__(f"libvapc_b200 error {code}: {msg}")
        self.code = code


class Grid(C.Structure):
    """vapc_grid_t (include/vapc_b200.h)."""

    _fields_ = [
        ("origin", C.c_double * 3),
        ("voxel_size", C.c_double),
        ("vmin", C.c_int64 * 3),
        ("bits", C.c_int32 * 3),
        ("key_bytes", C.c_int32),
    ]

    @property
    def total_bits(self):
        return int(self.bits[0] + self.bits[1] + self.bits[2])

    def copy(self):
        g = Grid()
        C.memmove(C.byref(g), C.byref(self), C.sizeof(Grid))
        return g


class LasTransform(C.Structure):
    """vapc_las_transform_t (include/vapc_b200.h): coordinate = X * scale + offset."""

    _fields_ = [("scale", C.c_double * 3), ("offset", C.c_double * 3)]


if not os.path.exists(LIB_PATH):
    raise ImportError(
        f"{LIB_PATH} not found: build it with vapc_b200/csrc/build.sh (nvcc, sm_100a). "
        "vapc_b200 has no CPU fallback."
    )

_lib = C.CDLL(LIB_PATH)

P = C.c_void_p
I64 = C.c_int64
I32 = C.c_int32
F64 = C.c_double
SZ = C.c_size_t
GP = C.POINTER(Grid)
TP = C.POINTER(LasTransform)

# name -> (restype, argtypes).  Order and meaning exactly as in include/vapc_b200.h.
SIGNATURES = {
    "vapc_abi_version": (C.c_int, []),
    "vapc_last_error": (C.c_char_p, []),
    "vapc_launch_count": (C.c_uint64, []),
    "vapc_set_l2_fetch_granularity": (C.c_int, [I32]),
    "vapc_point2voxel_set_ctas_per_sm": (C.c_int, [I32]),
    "vapc_profile_enable": (C.c_int, [I32]),
    "vapc_profile_collect": (C.c_int64, [C.c_char_p, I64]),
    "vapc_minmax_workspace_bytes": (SZ, []),
    "vapc_minmax_f64": (C.c_int, [P, P, P, I64, P, P, SZ, P]),
    "vapc_minmax_i32": (C.c_int, [P, P, P, I64, TP, P, P, SZ, P]),
    "vapc_las_coords_i32": (C.c_int, [P, P, P, I64, TP, P, P, P, P]),
    "vapc_pack_records_bucketed_i32": (C.c_int, [P, P, P, I64, TP, GP, P, P, P, P, C.POINTER(I32), C.POINTER(I32), P, P, SZ, P]),
    "vapc_voxel_coords_f64": (C.c_int, [P, P, P, I64, GP, P, P, P, P]),
    "vapc_grid_from_bounds": (C.c_int, [GP, C.POINTER(F64)]),
    "vapc_pack_keys_f64": (C.c_int, [P, P, P, I64, GP, P, P, P, P]),
    "vapc_unpack_keys": (C.c_int, [P, I64, GP, P, P, P, P]),
    "vapc_sort_workspace_bytes": (SZ, [I64]),
    "vapc_sort_debug_wide_lookback": (C.c_int, [I32]),
    "vapc_reduce_debug_lanes": (C.c_int, [I32]),
    "vapc_cov_reference_formula": (C.c_int, [P, P, P, I64, P, P]),
    "vapc_cog_reference_formula": (C.c_int, [P, P, P, I64, P, P]),
    "vapc_std_reference_formula": (C.c_int, [P, P, P, I64, P, P]),
    "vapc_sort_pairs": (C.c_int, [P, P, P, P, P, I32, I64, I32, I32, P, SZ, C.POINTER(I32), P]),
    "vapc_sort_pairs_bits": (C.c_int, [P, P, P, P, P, I32, I64, I32, I32, I32, P, SZ, C.POINTER(I32), P]),
    "vapc_scatter_u32": (C.c_int, [P, P, I64, P, P]),
    "vapc_segment_workspace_bytes": (SZ, [I64]),
    "vapc_count_voxels": (C.c_int, [P, I64, I32, P, P, SZ, P]),
    "vapc_reduce_moments": (C.c_int, [P, P, I64, GP, P, I64, P, P, P, P, P, P, P, P, P, SZ, P]),
    "vapc_record_indices": (C.c_int, [P, P, I64, P, P]),
    "vapc_bucket_workspace_bytes": (SZ, [I64]),
    "vapc_pack_records_bucketed": (C.c_int, [P, P, P, I64, GP, P, P, P, P, C.POINTER(I32), C.POINTER(I32), P, P, SZ, P]),
    "vapc_expand_bucket_keys": (C.c_int, [P, I64, P, I32, I32, P, P]),
    "vapc_sort_pairs_segmented": (C.c_int, [P, P, P, P, I32, I64, I32, I32, P, I32, P, SZ, C.POINTER(I32), P]),
    "vapc_sort_cells_segmented": (C.c_int, [P, I64, I32, I32, P, I32, P, P, P, P, P, SZ, P]),
    "vapc_point2voxel_table_bytes": (SZ, [I32]),
    "vapc_point2voxel_dense": (C.c_int, [P, I64, P, I64, I32, I32, P, SZ, P, P]),
    "vapc_voxel_stats": (C.c_int, [P, P, P, P, I64, GP, P, P, P, P, P, P, P, P]),
    "vapc_voxel_stats_axes": (C.c_int, [P, P, P, P, I64, GP, C.POINTER(I32), P, P, P, P, P, P, P, P]),
    "vapc_voxel_center_corner": (C.c_int, [P, I64, GP, P, P, P]),
    "vapc_closest_to_cog": (C.c_int, [P, P, I64, I32, P, P, I64, P, P, P, SZ, P]),
    "vapc_point_distance": (C.c_int, [P, P, P, I64, P, P, I64, P, P]),
    "vapc_gather_f64": (C.c_int, [P, P, I64, P, P]),
    "vapc_gather_u32": (C.c_int, [P, P, I64, P, P]),
    "vapc_mode_keys": (C.c_int, [P, P, I64, I32, I32, P, P, P]),
    "vapc_mode_count": (C.c_int, [P, I64, I32, I32, P, I64, F64, P, P, P]),
    "vapc_run_lengths": (C.c_int, [P, P, I64, I64, P, P, P, SZ, P]),
    "vapc_mode_count_runs": (C.c_int, [P, P, I64, I32, I64, F64, P, P, P]),
    "vapc_attr_stats": (C.c_int, [P, P, P, I64, P, P, P, P, P, P]),
    "vapc_f64_sort_keys": (C.c_int, [P, I64, P, P]),
    "vapc_attr_median": (C.c_int, [P, P, P, I64, P, P]),
    "vapc_pack_triples_i64": (C.c_int, [P, P, P, I64, GP, P, P, P]),
    "vapc_mask_table_bytes": (SZ, [I64, C.POINTER(I64)]),
    "vapc_mask_build": (C.c_int, [P, I64, P, I64, P]),
    "vapc_mask_lookup_points": (C.c_int, [P, P, P, I64, GP, P, I64, P, P]),
    "vapc_mask_lookup_keys": (C.c_int, [P, I64, P, I64, P, P]),
    "vapc_buffer_expand": (C.c_int, [P, P, P, I64, I32, P, P]),
    "vapc_join_sorted_keys": (C.c_int, [P, I64, P, I64, P, P, P]),
    "vapc_mahalanobis": (C.c_int, [P, P, I64, P, P, I64, P, P, I64, F64, P, P, P, P, P, P, P]),
    "vapc_cell_nn": (C.c_int, [P, P, P, I64, GP, P, I64, P, P, P, I64, I32, P, P, P]),
    "vapc_nn_all_pairs": (C.c_int, [P, P, P, P, I64, P, I64, P, P, P]),
    "vapc_cluster_workspace_bytes": (SZ, [I64]),
    "vapc_cluster_components": (C.c_int, [GP, P, I64, P, P, P, I64, F64, P, P, P, P, P, P, P, SZ, P]),
    "vapc_selftest_cell_nn_host": (C.c_int, [P, P, P, I64, GP, P, I64, P, P, P, I64, I32, P, P]),
    "vapc_selftest_cluster_host": (C.c_int, [GP, P, I64, P, P, P, I64, F64, P, P, P, P, P, P]),
    "vapc_merge_partials": (C.c_int, [P, P, I64, I32, P, P, P, I64, I64, P, P, P, P, P, SZ, P]),
    "vapc_plan_exchange": (C.c_int, [P, I32, I32, I32, P, P]),
    "vapc_merge_keys": (C.c_int, [P, I64, I64, P, I64, I32, P, P]),
    "vapc_selftest_plan_exchange_host": (C.c_int, [P, I32, I32, I32, P]),
    "vapc_global_keys": (C.c_int, [P, I64, GP, GP, I32, P, P, P]),
    "vapc_pack_partial_rows": (C.c_int, [P, P, P, P, I64, I64, I64, P, P]),
    "vapc_merge_partial_rows": (C.c_int, [P, P, I64, I32, P, P, P, P, I64, I64, I64, I64, P, P, P, P, P, SZ, P]),
    "vapc_merge_sorted_tables": (C.c_int, [P, P, P, P, I64, I64, I64, P, P, P, P, I64, P, I32, I64, P, P, P, P, P]),
    "vapc_merge_sorted_tables_stats": (C.c_int, [P, P, P, P, I64, I64, I64, P, P, P, P, I64, P, I32, I64, P, P, P, P, GP, C.POINTER(I32), P, P, P, P, P, P, P, P]),
    "vapc_reduce_candidates": (C.c_int, [P, P, P, I64, I64, P, P, P]),
    "vapc_key_histogram": (C.c_int, [P, I64, I32, I32, I32, P, P]),
    "vapc_laz_decode": (C.c_int, [P, I64, I64, I64, I32, I32, C.c_uint32, P, I32, P]),
    "vapc_laz_encode": (C.c_int, [P, I64, I32, C.c_uint32, P, I32, I64, P, I64, C.POINTER(I64)]),
    "vapc_selftest_eig3_host": (C.c_int, [P, I64, P]),
    "vapc_selftest_chan_host": (C.c_int, [P, I64, P]),
    "vapc_selftest_mahalanobis_host": (C.c_int, [P, P, I64, P, P]),
}

for _name, (_res, _args) in SIGNATURES.items():
    _fn = getattr(_lib, _name)
    _fn.restype = _res
    _fn.argtypes = _args


def last_error() -> str:
    return _lib.vapc_last_error().decode("utf-8", "replace")


def check(code: int):
    """Map a C status to the Python exception the reference would raise at that point:
    bad arguments -> ValueError (vapc.py:66-69), everything else -> VapcError."""
    if code == VAPC_OK:
        return
    msg = last_error()
    if code == VAPC_E_INVALID:
        raise ValueError(msg)
    raise VapcError(code, msg)


# optional observer(name, phase) with phase in {"before", "after"}: bench.py uses it to bracket every
# C-ABI call with CUDA events on the launching stream
observer = None


def call(name: str, *args):
    """Call a status-returning entry point and raise on failure."""
    obs = observer
    if obs is not None:
        obs(name, "before")
    code = getattr(_lib, name)(*args)
    if obs is not None:
        obs(name, "after")
    check(code)


def profile_collect():
    """{kernel: (launches, total_ms)} of the launches recorded since vapc_profile_enable(1) / the last collect."""
    buf = C.create_string_buffer(1 << 16)
    _lib.vapc_profile_collect(buf, len(buf))
    out = {}
    for ln in buf.value.decode().splitlines():
        name, cnt, ms = ln.split("\t")
        out[name] = (int(cnt), float(ms))
    return out


def raw(name: str):
    return getattr(_lib, name)
