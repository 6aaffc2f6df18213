"""`Vapc` — drop-in for the reference's `vapc.Vapc` (src/vapc/vapc.py:8-1357) on top of the
B200 engine.

Same constructor, attributes, state flags, method names and DataFrame column names as the
reference; `self.df` is a host pandas DataFrame whenever a caller looks at it (callers read and
assign it directly, tests/test_voxel.py:82-83) while every per-point computation runs on the GPU
through `vapc_b200.engine` (C ABI calls).

The reference materialises every per-voxel value per point (16+ float64 columns x N rows).
Here a computed column stays on the device as its [V] voxel column + the point -> voxel map until somebody
reads `.df`: the `df` property then broadcasts the pending columns into the frame, in the reference's
column order.  `compute_requested_attributes()` followed by `reduce_to_voxels()` — what the command-line
tools do — never builds the N-row columns at all: the reduced frame is formed from the V-row table.
A caller that has been handed the frame may edit it in place, so every read of `.df` also drops the device
copy of the cloud; it is rebuilt from the frame's X, Y, Z on the next computation.

Host-side work is limited to what the DataFrame container itself requires (adding / dropping /
reordering columns, row selection with a mask computed on the GPU, the voxel MultiIndex).
"""
from __future__ import annotations

import threading
import warnings
from concurrent.futures import ThreadPoolExecutor

import numpy as np
import pandas as pd
import torch

from . import engine as E
from .utilities import timeit, trace

_STAGING = {}  # device index -> pinned buffers: multi-buffered device -> host staging shared by all Vapc objects
_STAGING_LOCK = threading.Lock()
_STAGING_BUFFERS = 4  # the landing copies (pinned -> fresh pageable memory, i.e. page faults) run at ~10 GB/s per thread: four in flight


def _download_rows(block: np.ndarray, rows, make):
    """block[i] = make(i) (a float64 device column) for every i in rows.  Device -> pinned staging copies rotate through a few
    buffers while worker threads move the previous columns into the (pageable) block: the block is filled at the speed of the host
    memcpy instead of a pageable cudaMemcpy.  The module-wide staging buffers are used under a lock (one download at a time)."""
    if not len(rows):
        return
    width = block.shape[1]
    dev = torch.cuda.current_device()
    with _STAGING_LOCK:
        bufs = _STAGING.get(dev)
        if bufs is None or bufs[0].numel() < width:
            bufs = [torch.empty(max(width, 1), dtype=torch.float64).pin_memory() for _ in range(_STAGING_BUFFERS)]
            _STAGING[dev] = bufs

        def land(ev, buf, i):
            ev.synchronize()
            np.copyto(block[i], buf.numpy()[:width])

        with ThreadPoolExecutor(max_workers=_STAGING_BUFFERS) as pool:
            pending = [None] * _STAGING_BUFFERS
            for j, i in enumerate(rows):
                k = j % _STAGING_BUFFERS
                if pending[k] is not None:
                    pending[k].result()
                t = make(i)
                bufs[k][:width].copy_(t.to(torch.float64), non_blocking=True)
                ev = torch.cuda.Event()
                ev.record()
                pending[k] = pool.submit(land, ev, bufs[k], i)
                del t
            for f in pending:
                if f is not None:
                    f.result()


_COV_NAMES = ["cov_xx", "cov_xy", "cov_xz", "cov_yx", "cov_yy", "cov_yz", "cov_zx", "cov_zy", "cov_zz"]
_COV9_FROM6 = [0, 1, 2, 1, 3, 4, 2, 4, 5]


def _u32(t: torch.Tensor) -> torch.Tensor:
    """uint32 values stored in an int32 tensor -> int64."""
    return t.to(torch.int64) & 0xFFFFFFFF


class Vapc:
    AVAILABLE_COMPUTATIONS = [
        "voxel_index",
        "point_count",
        "point_density",
        "percentage_occupied",
        "covariance_matrix",
        "eigenvalues",
        "geometric_features",
        "center_of_gravity",
        "distance_to_center_of_gravity",
        "std_of_cog",
        "closest_to_center_of_gravity",
        "center_of_voxel",
        "corner_of_voxel",
    ]

    def __init__(self, voxel_size: float, origin: list = None, attributes: dict = None, compute: list = None,
                 return_at: str = "closest_to_center_of_gravity"):
        """Same parameters and validation as the reference constructor (vapc.py:26-110)."""
        self.voxel_size = voxel_size
        self.origin = origin if origin is not None else [0, 0, 0]
        self.attributes = attributes if attributes is not None else {}
        self.compute = compute if compute is not None else []
        self.return_at = return_at
        if self.voxel_size <= 0:
            raise ValueError("voxel_size must be a positive number.")
        if not isinstance(self.origin, list) or len(self.origin) != 3:
            raise ValueError("origin must be a list of three coordinates [x, y, z].")
        self.AVAILABLE_COMPUTATIONS = list(Vapc.AVAILABLE_COMPUTATIONS)
        # not in the reference: which numbers the cov_* columns carry.  "exact" = M2 / (n - 1) from centred moments (~1e-13 of the true
        # covariance at any coordinate offset); "reference" = the reference's raw-moment columns bit for bit (vapc.py:947-1006), which
        # lose all digits at UTM-sized coordinates.  Eigenvalues and features never depend on it (the reference takes them from np.cov).
        self.covariance_formula = "exact"
        # not in the reference either: which double the centroid is.  "reference" (default) = pandas' groupby mean bit for bit (Kahan sum
        # in row order / count, vapc.py:843-850): cog_*, distance, min_distance and the rows kept by the closest-to-centroid reduction
        # (exact ties included) then EQUAL the reference's; "exact" = the mean from the reduction's shifted moments (no extra pass; equal
        # to within a bit, ties may fall differently).  Nothing is lost either way: both are correctly rounded to the last bit or two.
        self.centroid_formula = "reference"
        # the std_* columns: "exact" = sqrt(M2 / (n - 1)) from the shifted moments; "reference" = pandas' groupby std bit for bit (Welford in
        # row order, vapc.py:873-880), which is ~1e-7 off at UTM-sized coordinates — for users who diff against the reference
        self.std_formula = "exact"
        self._df = None      # the host frame: original columns + whatever has been materialised
        self._lazy = {}      # column name -> (kind, device tensor, cloud); kind "voxel": [V] column to broadcast, "point": [N] column
        self._cols = None    # logical column order (frame columns + pending ones), None = the frame's own
        self._exposed = True  # a caller holds the frame (and may have edited it) since the cloud was built
        self._pinned = {}    # (dtype, device index) -> reusable pinned staging buffer of THIS object
        self.original_attributes = None
        # state flags of the lazy-dependency protocol (vapc.py:90-110)
        self.attributes_up_to_data = False
        self.voxelized = False
        self.voxel_index = False
        self.point_count = False
        self.point_density = False
        self.percentage_occupied = False
        self.covariance_matrix = False
        self.eigenvalues = False
        self.geometric_features = False
        self.center_of_gravity = False
        self.distance_to_center_of_gravity = False
        self.std_of_cog = False
        self.closest_to_center_of_gravity = False
        self.center_of_voxel = False
        self.corner_of_voxel = False
        self.attributes_per_voxel = False
        self.drop_columns = []
        self.new_column_names = {}
        self.reduced = False
        self.reduction_point = None
        self.offset_applied = False
        # closest-to-centroid reduction: "all" keeps every exactly tied point like the reference
        # (vapc.py:1195); "first" keeps the lowest original index only (north_star tie-break).
        self.closest_ties = "all"
        self._cloud = None
        self._cloud_sig = None

    # ------------------------------------------------------------------------------------
    # the frame: host columns + columns still pending on the device
    # ------------------------------------------------------------------------------------
    @property
    def df(self):
        """The per-point (or reduced) pandas DataFrame, every computed column present.  Reading it hands the frame to the caller:
        pending device columns are broadcast into it first, and the device copy of the cloud is dropped afterwards (the caller may
        edit the frame in place; the next computation re-reads X, Y, Z)."""
        if self._df is not None:
            self._materialise()
            self._exposed = True
        return self._df

    @df.setter
    def df(self, value):
        self._df = value
        self._lazy = {}
        self._cols = None
        self._cloud = None
        self._exposed = True

    def _columns(self):
        return list(self._df.columns) if self._cols is None else list(self._cols)

    def _has(self, name):
        return name in self._lazy or (self._df is not None and name in self._df.columns)

    def _add_lazy(self, name, kind, tensor, cloud):
        if self._cols is None:
            self._cols = list(self._df.columns)
        if name not in self._cols:
            self._cols.append(name)
        self._lazy[name] = (kind, tensor, cloud)

    def _lazy_device(self, name) -> torch.Tensor:
        """A pending column as an [N] device tensor (the voxel column broadcast through the point -> voxel map)."""
        kind, t, cloud = self._lazy[name]
        if kind == "voxel":
            if t.dtype in (torch.float64, torch.int32):
                t = cloud.broadcast(t.contiguous())
            else:
                t = t[_u32(cloud.point2voxel)]
        return t

    def _lazy_host(self, name) -> np.ndarray:
        """A pending column as a fresh host array of N rows."""
        return self._host(self._lazy_device(name)).copy()

    def _materialise(self, only=None):
        """Broadcast pending columns (all, or the named ones) into the host frame, keeping the logical column order.  The 8-byte
        columns (float64, and int64 carried as bit patterns) land in two [k, N] blocks through the double-buffered pinned staging of
        _download_rows; when everything pending is materialised the frame is rebuilt from column views in the final order, so that
        neither pandas' consolidation nor a re-ordering copy touches the N x 39 values again."""
        if not self._lazy:
            if self._cols is not None and list(self._df.columns) == self._cols:
                self._cols = None
            return
        names = [c for c in self._cols if c in self._lazy and (only is None or c in only)]
        if not names:
            return
        n = len(self._df)
        new = {}
        for dt_t, dt_n in ((torch.float64, np.float64), (torch.int64, np.int64)):
            group = [c for c in names if self._lazy[c][1].dtype == dt_t]
            if group and n:
                block = np.empty((len(group), n), dtype=dt_n)
                _download_rows(block.view(np.float64), list(range(len(group))), lambda i: self._lazy_device(group[i]).view(torch.float64))
                new.update({c: block[i] for i, c in enumerate(group)})
        for c in names:
            if c not in new:
                new[c] = self._lazy_host(c)
        for c in names:
            del self._lazy[c]
        plain = all(isinstance(self._df[c].dtype, np.dtype) for c in self._df.columns)
        if not self._lazy and plain:
            # final order in one go: every column its own array (existing ones as views), no block consolidation, no reorder copy
            def existing(c):
                a = self._df[c].to_numpy()
                if not a.flags.writeable:  # pandas hands out read-only views (copy-on-write); the new frame takes the column over
                    try:
                        a.flags.writeable = True
                    except ValueError:
                        a = a.copy()
                return a

            cols = {c: (new[c] if c in new else existing(c)) for c in self._cols}
            self._df = pd.DataFrame(cols, index=self._df.index, copy=False)
            self._cols = None
        else:
            fresh = [c for c in names if c not in self._df.columns]
            for c in names:
                if c not in fresh:
                    self._df[c] = new[c]
            if fresh:
                add = pd.DataFrame({c: new[c] for c in fresh}, index=self._df.index, copy=False)
                self._df = pd.concat([self._df, add], axis=1)
            if not self._lazy:
                if list(self._df.columns) != self._cols:
                    self._df = self._df[self._cols]
                self._cols = None
        self._resign()  # same points, possibly a new frame object

    def _host_column(self, name, dtype=None) -> np.ndarray:
        """Column `name` as a host array whether it is pending or in the frame."""
        if name in self._lazy:
            a = self._lazy_host(name)
        else:
            a = self._df[name].to_numpy()
        return a if dtype is None else np.asarray(a, dtype=dtype)

    def _assign_host(self, name, array):
        """A host-computed column into the frame (and into the logical order while other columns are pending)."""
        self._lazy.pop(name, None)
        self._df[name] = array
        if self._cols is not None and name not in self._cols:
            self._cols.append(name)

    def _drop_device_state(self):
        """The frame's rows or coordinates changed: pending columns and the device cloud no longer describe it."""
        self._lazy = {}
        self._cols = None
        self._cloud = None

    # ------------------------------------------------------------------------------------
    # device cloud cache
    # ------------------------------------------------------------------------------------
    def _signature(self):
        """What the device cloud was built from: the frame object, its length, voxel size, origin, and a fingerprint of X / Y / Z at
        up to 1024 evenly spaced rows per axis (a few microseconds: it catches a column that was shifted, scaled or replaced through a
        reference the caller kept to the frame they assigned).  An edit of single cells through such a reference is not seen — the
        frame read from `.df` always is (`_exposed`) —: `invalidate()` states it explicitly."""
        df = self._df
        n = len(df)
        rows = np.linspace(0, n - 1, min(n, 1024)).astype(np.int64) if n else None
        probe = tuple(hash(df[c].to_numpy()[rows].tobytes()) if n and c in df.columns else None for c in ("X", "Y", "Z"))
        return (id(df), n, float(self.voxel_size), tuple(float(o) for o in self.origin), probe)

    def invalidate(self):
        """Forget the device copy of the cloud: call after editing, in place, a frame this object already computed on through a
        reference of your own (pending columns are written into the frame first; the next compute_* re-reads X, Y, Z)."""
        if self._df is not None:
            self._materialise()
        self._drop_device_state()
        self._exposed = True

    def _get_cloud(self) -> E.VoxelCloud:
        """The device-side voxelisation of the frame's X, Y, Z columns.  Rebuilt when the frame was handed out since the last
        build (the caller may have edited it), was replaced, or voxel_size / origin changed."""
        if self._df is None:
            raise AttributeError("no point cloud loaded: assign `df` or call get_data_from_data_handler()")
        sig = self._signature()
        if self._cloud is None or self._exposed or sig != self._cloud_sig:
            self._materialise()  # pending columns belong to the old cloud
            given = None
            if self.voxelized and all(c in self._df.columns for c in ("voxel_x", "voxel_y", "voxel_z")):
                # the frame carries voxel columns (handed out, assigned or edited by the caller): the reference groups by THOSE,
                # whatever the coordinates say now (vapc.py:724, :843, :959)
                given = [self._df[c].to_numpy(np.int64) for c in ("voxel_x", "voxel_y", "voxel_z")]
            self._cloud = E.VoxelCloud.build(self._df["X"].to_numpy(np.float64), self._df["Y"].to_numpy(np.float64),
                                             self._df["Z"].to_numpy(np.float64), float(self.voxel_size), self.origin, voxels=given)
            self._cloud_sig = self._signature()
            self._exposed = False
            self._voxelized_cloud = self._cloud if given is not None else None
            self._cloud_given = given is not None
        self._cloud.centroid_formula = self.centroid_formula  # drops a cached centroid of the other kind
        return self._cloud

    def _reset_index(self):
        """The reference's merges leave the frame with a fresh 0..n-1 index.  A frame that has it already is left alone:
        reset_index would copy every column of it (5 x 80 MB at 1e7 points: 0.09 s of a 0.35 s pipeline)."""
        idx = self._df.index
        if isinstance(idx, pd.RangeIndex) and idx.start == 0 and idx.step == 1:
            return
        self._df = self._df.reset_index(drop=True)
        self._resign()  # same points in a new frame object: keep the device cloud

    def _resign(self):
        """Re-take the cache signature after this class itself replaced the frame object (reset_index, reordering)."""
        if self._cloud is not None and self._cloud_sig is not None:
            self._cloud_sig = (id(self._df), len(self._df)) + tuple(self._cloud_sig[2:])  # grid parameters as they were at build time

    def _set_per_point(self, name: str, voxel_column: torch.Tensor):
        """Column `name` = the [V] voxel column broadcast to the points (pending until `.df` is read)."""
        self._add_lazy(name, "voxel", voxel_column, self._get_cloud())

    def _set_col(self, name: str, t: torch.Tensor):
        """Column `name` = an [N] device column (pending until `.df` is read)."""
        self._add_lazy(name, "point", t, self._get_cloud())

    def _host(self, t: torch.Tensor) -> np.ndarray:
        """Device column -> host array through a reusable pinned staging buffer of this object (a pageable copy of an 80 MB column
        runs at ~2.5 GB/s, the pinned one at PCIe speed).  The result is a view of the staging buffer: valid until the next call."""
        key = (t.dtype, t.device.index)
        buf = self._pinned.get(key)
        if buf is None or buf.numel() < t.numel():
            buf = torch.empty(max(t.numel(), 1), dtype=t.dtype).pin_memory()
            self._pinned[key] = buf
        view = buf[: t.numel()]
        view.copy_(t.contiguous(), non_blocking=False)
        return view.numpy()

    def release(self):
        """Free the pinned staging buffers and the device copy of the cloud (they are re-created on demand)."""
        self._materialise() if self._df is not None else None
        self._pinned = {}
        self._cloud = None

    def _validate_compute_list(self):
        invalid = [c for c in self.compute if c not in self.AVAILABLE_COMPUTATIONS]
        if invalid:
            raise ValueError(f"Invalid computation(s) requested: {invalid}. "
                             f"Available computations are: {self.AVAILABLE_COMPUTATIONS}")

    # ------------------------------------------------------------------------------------
    @trace
    @timeit
    def get_data_from_data_handler(self, data_handler):
        """Moves the DataFrame out of the data handler (vapc.py:132-149)."""
        if data_handler.df is None:
            raise AttributeError("The provided data_handler does not have a 'df' attribute.")
        self.df = data_handler.df
        data_handler.df = None
        self.original_attributes = self._df.columns.tolist()

    @trace
    @timeit
    def compute_reduction_point(self):
        """[int(min X), int(min Y), int(min Z)] (vapc.py:153-162); the minima come from the GPU bbox kernel."""
        dev = torch.device("cuda", torch.cuda.current_device())
        E._require_cuda()
        cols = [E._dev_f64(self._df[c].to_numpy(np.float64), dev) for c in ("X", "Y", "Z")]
        mm = E.coordinate_bounds(*cols)
        self.reduction_point = [int(mm[0]), int(mm[1]), int(mm[2])]

    @trace
    @timeit
    def compute_offset(self):
        """Toggle the subtraction of `reduction_point` from X, Y, Z (vapc.py:166-184)."""
        if self.reduction_point is None:
            self.compute_reduction_point()
        sign = 1 if self.offset_applied else -1
        self._materialise()
        for c, r in zip(("X", "Y", "Z"), self.reduction_point):
            self._df[c] = self._df[c] + sign * r
        self._drop_device_state()  # the coordinates changed
        self.offset_applied = not self.offset_applied

    def _voxelize(self):
        """voxel_x / voxel_y / voxel_z = floor((coord - origin) / voxel_size) as int64 (vapc.py:188-207): the voxel table's triples
        broadcast to the points (pending until `.df` is read)."""
        self.voxelized = False  # (re-)voxelise from the coordinates, not from voxel columns the frame may carry
        if getattr(self, "_cloud_given", False):
            self._cloud = None
        cloud = self._get_cloud()
        for name, col in zip(("voxel_x", "voxel_y", "voxel_z"), cloud.voxel_coords()):
            self._add_lazy(name, "voxel", col, cloud)
        self._voxelized_cloud = cloud
        self.voxelized = True

    @trace
    @timeit
    def voxelize(self):
        """voxel_x / voxel_y / voxel_z = floor((coord - origin) / voxel_size) as int64; returns the frame like the reference
        (vapc.py:188-207)."""
        self._voxelize()
        return self.df

    @trace
    @timeit
    def compute_requested_attributes(self):
        """Dispatcher over `self.compute` (vapc.py:211-252)."""
        self._validate_compute_list()
        for name in self.compute:
            if getattr(self, name) is True:
                continue
            method = getattr(self, f"compute_{name}", None)
            if callable(method):
                try:
                    method()
                except Exception as e:
                    print(f"Error computing '{name}': {e}")
                    raise e
            else:
                print(f"Invalid computation name: '{name}'")
                raise ValueError(f"Invalid computation name: '{name}'")

    def update_attribute_dictionary(self, remove_cols=None):
        """Default statistic 'mean' for every original attribute unless specified (vapc.py:254-299)."""
        if remove_cols is None:
            remove_cols = ["X", "Y", "Z", "bit_fields", "raw_classification", "scan_angle_rank", "user_data", "point_source_id"]
        original = [c for c in self.original_attributes if c not in remove_cols]
        attributes = {c: "mean" for c in original}
        attributes.update(self.attributes)
        self.attributes = attributes
        self.attributes_up_to_data = True

    # ------------------------------------------------------------------------------------
    @trace
    @timeit
    def compute_requested_statistics_per_attributes(self):
        """Per-voxel statistics of attribute columns broadcast to the points, column f"{attr}_{stat}"
        (vapc.py:303-472).  `mode` and `mode_count,<p>` (vapc.py:407-433) run on the GPU as a
        sort by (voxel, value) + run lengths; mean / sum / min / max / std / var (ddof = 0) / median are
        segmented reductions over the voxel-sorted attribute."""
        if self.attributes == {}:
            print("No attributes specified for computation. Skipping.")
            return
        if not self.voxelized:
            self._voxelize()
            self.drop_columns += ["voxel_x", "voxel_y", "voxel_z"]
        if not self.attributes_up_to_data:
            self.update_attribute_dictionary()
        required = ["voxel_x", "voxel_y", "voxel_z"] + list(self.attributes.keys())
        missing = [c for c in required if not self._has(c)]
        if missing:
            raise KeyError(f"Missing columns in `self.df`: {missing}")
        cloud = self._get_cloud()
        new_cols = {}
        for attr, stats_list in self.attributes.items():
            if not isinstance(stats_list, list):
                stats_list = [stats_list]
            values = self._host_column(attr, np.float64)
            thresholds, want_mode = [], False
            for stat in stats_list:
                if stat == "mode":
                    want_mode = True
                elif "mode_count" in stat:
                    try:
                        _, p = stat.split(",")
                        thresholds.append((stat, float(p)))
                    except ValueError as exc:
                        raise ValueError(f"Invalid 'mode_count' specification for attribute '{attr}': '{stat}'") from exc
            mode_res, count_res = (None, {})
            if want_mode or thresholds:
                try:
                    mode_res, count_res = cloud.mode_count(values, [p for _, p in thresholds], want_mode=True)
                except ValueError as exc:
                    if thresholds:
                        raise ValueError(f"Cannot compute 'mode_count' for continuous attribute '{attr}'") from exc
                    raise
            simple = None
            for stat in stats_list:
                col = f"{attr}_{stat}"
                if stat == "mode":
                    new_cols[col] = mode_res
                elif "mode_count" in stat:
                    new_cols[col] = count_res[dict(thresholds)[stat]]
                elif stat in ("mean", "std", "var", "median", "min", "max", "sum"):
                    if simple is None:
                        simple = E.attribute_statistics(cloud, values, median="median" in stats_list)
                    new_cols[col] = simple[stat]
                else:
                    print(f"Unknown aggregation type '{stat}' for attribute '{attr}'. Skipping.")
        # the reference merges the aggregated frame back with how="left" (index reset, vapc.py:468-470)
        self._reset_index()
        for col, table in new_cols.items():
            self._add_lazy(col, "voxel", table, cloud)
        self.attributes_per_voxel = True

    # ------------------------------------------------------------------------------------
    @trace
    @timeit
    def compute_voxel_index(self):
        """3-level MultiIndex idx_voxel_{x,y,z} equal to the voxel triple, columns kept (vapc.py:476-491)."""
        if self.voxel_index:
            return
        if not self.voxelized:
            self._voxelize()
            self.drop_columns += ["voxel_x", "voxel_y", "voxel_z"]
        self._materialise(only=("voxel_x", "voxel_y", "voxel_z"))  # the MultiIndex lives in the frame
        self._df.set_index(["voxel_x", "voxel_y", "voxel_z"], inplace=True, drop=False)
        self._df.index.set_names(["idx_voxel_x", "idx_voxel_y", "idx_voxel_z"], inplace=True)
        self.voxel_index = True
        self._resign()

    @trace
    @timeit
    def compute_voxel_buffer(self, buffer_size: int = 1):
        """Unique voxels of this cloud, each expanded by all (2b+1)^3 offsets and de-duplicated in
        first-occurrence order; overwrites `self.df` with that mask frame and returns it
        (vapc.py:495-541, including its X/Y/Z = (v + vs/2) * vs quirk)."""
        if not self.voxelized:
            self._voxelize()
        if all(c in self._lazy for c in ("voxel_x", "voxel_y", "voxel_z")):
            # the voxel table IS the unique triples; first-occurrence order = ascending first point index
            cloud = self._lazy["voxel_x"][2]
            order = torch.argsort(_u32(cloud.first_idx))
            uniq = [c[order] for c in cloud.voxel_coords()]
        else:
            uniq = E.unique_voxels_first_occurrence(*[self._host_column(c, np.int64) for c in ("voxel_x", "voxel_y", "voxel_z")])
        expanded = E.buffer_expand(uniq[0], uniq[1], uniq[2], int(buffer_size))
        buf = E.unique_voxels_first_occurrence(expanded[0], expanded[1], expanded[2])
        self.df = pd.DataFrame({n: c.cpu().numpy() for n, c in zip(("voxel_x", "voxel_y", "voxel_z"), buf)})
        self.voxelized = True
        self.voxel_index = False
        self.compute_voxel_index()
        self._df["Y"] = (self._df["voxel_y"] + self.voxel_size / 2) * self.voxel_size
        self._df["Z"] = (self._df["voxel_z"] + self.voxel_size / 2) * self.voxel_size
        self._df["X"] = (self._df["voxel_x"] + self.voxel_size / 2) * self.voxel_size
        self.buffer_df = self._df
        self._drop_device_state()
        return self._df

    def _cloud_if_current(self):
        """The device cloud if it still describes the frame (not handed out, not replaced, same grid), else None."""
        if self._cloud is None or self._exposed or self._df is None or self._signature() != self._cloud_sig:
            return None
        self._cloud.centroid_formula = self.centroid_formula
        return self._cloud

    def _own_voxel_cloud(self):
        """The device cloud if the frame's voxel_x / voxel_y / voxel_z are its own voxelisation (set by this class, frame not
        handed out since): per-voxel answers can then be formed on the V-row table and broadcast."""
        cloud = self._cloud_if_current()
        return cloud if cloud is not None and getattr(self, "_voxelized_cloud", None) is cloud else None

    def _mask_voxels(self, vapc_mask):
        if self._df is None or vapc_mask._df is None:
            raise AttributeError("Both `self.df` and `vapc_mask.df` must exist.")
        if self.voxelized is False:
            self._voxelize()
            self.drop_columns += ["voxel_x", "voxel_y", "voxel_z"]
        if vapc_mask.voxelized is False:
            vapc_mask._voxelize()
        if vapc_mask.voxel_index is False:
            vapc_mask.compute_voxel_index()
        if self.voxel_index is False:
            self.compute_voxel_index()
        idx = vapc_mask._df.index
        return [np.asarray(idx.get_level_values(k), dtype=np.int64) for k in range(3)]

    def _per_point_of_voxels(self, cloud, voxel_values: torch.Tensor) -> np.ndarray:
        """[V] int32 device values -> host array of N rows (broadcast through the point -> voxel map)."""
        return self._host(cloud.broadcast(voxel_values.to(torch.int32).contiguous())).copy()

    @trace
    @timeit
    def select_by_mask(self, vapc_mask, segment_in_or_out="in"):
        """Keep ("in") or remove ("out") the points whose voxel triple is in the mask's voxel set
        (vapc.py:545-609).  Membership is a GPU hash-set probe — of the V voxel rows when the voxel columns are this object's own
        voxelisation, else of the frame's index; original order is kept, the voxel columns are dropped afterwards and
        `voxelized` is reset, as in the reference."""
        mv = self._mask_voxels(vapc_mask)
        if segment_in_or_out not in ("in", "out"):
            raise ValueError("Parameter 'segment_in_or_out' must be either 'in' or 'out'.")
        mask = E.VoxelMask(mv[0], mv[1], mv[2], float(self.voxel_size), self.origin)
        cloud = self._own_voxel_cloud()
        if cloud is not None:
            member = self._per_point_of_voxels(cloud, mask.contains_voxels(*cloud.voxel_coords())).astype(bool)
        else:
            own = [np.asarray(self._df.index.get_level_values(k), dtype=np.int64) for k in range(3)]
            member = mask.contains_voxels(*own).cpu().numpy().astype(bool)
        self._materialise()
        frame = self._df.loc[member if segment_in_or_out == "in" else ~member]
        for attr in ["voxel_x", "voxel_y", "voxel_z"]:
            if attr in frame.columns:
                frame = frame.drop([attr], axis=1)
        self._df = frame
        self.voxelized = False
        self._drop_device_state()

    @trace
    @timeit
    def label_by_mask(self, vapc_mask, label_attributes=["voxel_x", "voxel_y", "voxel_z"]):
        """Left-join selected mask columns onto the points by voxel triple (vapc.py:614-677).  With a mask of unique voxels the
        join runs on the GPU: sorted-key join of the two voxel tables -> mask row of every voxel -> broadcast to the points; the
        mask's columns are then gathered by that row (NaN where the voxel is not in the mask, as the left join gives).  A mask
        with repeated voxels multiplies rows in the reference: that case goes through pandas' own join."""
        self._mask_voxels(vapc_mask)
        mask_df = vapc_mask._df
        common = [c for c in label_attributes if c in mask_df.columns]
        self._materialise(only=("voxel_x", "voxel_y", "voxel_z"))
        if common:
            overlap = [c for c in common if self._has(c)]
            if overlap:  # pandas' join refuses equal column names without a suffix (the reference's default arguments end here)
                raise ValueError(f"columns overlap but no suffix specified: {pd.Index(overlap)}")
            cloud = self._own_voxel_cloud()
            if cloud is not None and mask_df.index.is_unique:
                mv = [np.asarray(mask_df.index.get_level_values(k), dtype=np.int64) for k in range(3)]
                dev = cloud.device
                i1, i2, _, _ = E.join_voxel_tables(cloud.voxel_coords(), [torch.from_numpy(a).to(dev) for a in mv])
                row_of_voxel = torch.full((cloud.nvoxels,), -1, dtype=torch.int32, device=dev)
                row_of_voxel[i1] = i2.to(torch.int32)
                mrow = self._per_point_of_voxels(cloud, row_of_voxel).astype(np.int64)
                found = mrow >= 0
                taken = mask_df[common].take(np.where(found, mrow, 0))
                if not found.all():
                    taken = taken.where(np.broadcast_to(found[:, None], taken.shape))  # NaN (and the dtype promotion of the join)
                self._materialise()
                for c in common:
                    self._df[c] = taken[c].to_numpy()
            else:
                self._materialise()
                self._df = self._df.join(mask_df[common], how="left")
        self._materialise()
        for attr in ["voxel_x", "voxel_y", "voxel_z"]:
            if attr in self._df.columns:
                self._df = self._df.drop([attr], axis=1)
        self.voxelized = False
        self._drop_device_state()

    # ------------------------------------------------------------------------------------
    @trace
    @timeit
    def compute_point_count(self):
        """point_count = points in the point's voxel, int64 (vapc.py:715-726)."""
        if self.voxelized is False:
            self._voxelize()
        cloud = self._get_cloud()
        self._set_per_point("point_count", _u32(cloud.count))
        self.point_count = True

    @trace
    @timeit
    def compute_point_density(self):
        """point_count / voxel_size**3 (vapc.py:730-741)."""
        if self.point_count is False:
            self.compute_point_count()
            if "point_count" not in self.compute:
                self.drop_columns += ["point_count"]
        cloud = self._get_cloud()
        self._set_per_point("point_density", cloud.stats(density=True).density)
        self.point_density = True

    @trace
    @timeit
    def compute_percentage_occupied(self):
        """Occupied share of the voxel bounding box, stored in `self.percentage_occupied` (vapc.py:745-773)."""
        if self.voxel_index is False:
            self.compute_voxel_index()
        self.percentage_occupied = self._get_cloud().percentage_occupied()

    def compute_distance_to_center_of_gravity(self):
        """distance = |point - voxel centroid| per point (vapc.py:776-795)."""
        if self.center_of_gravity is False:
            self.compute_center_of_gravity()
            if "center_of_gravity" not in self.compute:
                self.drop_columns += ["cog_x", "cog_y", "cog_z"]
        self._set_col("distance", self._get_cloud().point_distance())
        self.distance_to_center_of_gravity = True

    @trace
    @timeit
    def compute_closest_to_center_of_gravity(self):
        """min_distance = smallest distance in the point's voxel (vapc.py:799-825; the reference's
        merge resets the index)."""
        if not self.center_of_gravity:
            self.compute_center_of_gravity()
            if "center_of_gravity" not in self.compute:
                self.drop_columns += ["cog_x", "cog_y", "cog_z"]
        if not self.distance_to_center_of_gravity:
            self.compute_distance_to_center_of_gravity()
            if "distance_to_center_of_gravity" not in self.compute:
                self.drop_columns += ["distance"]
        cloud = self._get_cloud()
        min_dist, argmin = cloud.closest_to_cog()
        self._closest_argmin = _u32(argmin)
        self._reset_index()
        self._set_per_point("min_distance", min_dist)
        self.closest_to_center_of_gravity = True

    @trace
    @timeit
    def compute_center_of_gravity(self):
        """cog_x / cog_y / cog_z = per-voxel mean of X, Y, Z (vapc.py:829-852)."""
        if not self.voxelized:
            self._voxelize()
            self.drop_columns += ["voxel_x", "voxel_y", "voxel_z"]
        cloud = self._get_cloud()
        cog = cloud.stats(cog=True).cog
        for k, name in enumerate(("cog_x", "cog_y", "cog_z")):
            self._set_per_point(name, cog[k])
        self.center_of_gravity = True

    @trace
    @timeit
    def compute_std_of_cog(self):
        """std_x / std_y / std_z = per-voxel sample standard deviation, ddof = 1 (vapc.py:856-887)."""
        if not self.voxelized:
            self._voxelize()
            self.drop_columns += ["voxel_x", "voxel_y", "voxel_z"]
        cloud = self._get_cloud()
        if self.std_formula == "reference":
            std = cloud.std_reference_formula()  # pandas' Welford recurrence in row order, bit for bit (for diffing)
        elif self.std_formula == "exact":
            std = cloud.stats(std=True).std
        else:
            raise ValueError("std_formula must be 'exact' or 'reference'")
        for k, name in enumerate(("std_x", "std_y", "std_z")):
            self._set_per_point(name, std[k])
        self.std_of_cog = True

    @trace
    @timeit
    def compute_center_of_voxel(self):
        """center_* = ((voxel * vs) + vs / 2) + origin (vapc.py:891-907)."""
        if self.voxelized is False:
            self._voxelize()
            self.drop_columns += ["voxel_x", "voxel_y", "voxel_z"]
        center, _ = self._get_cloud().center_corner(center=True, corner=False)
        for k, name in enumerate(("center_x", "center_y", "center_z")):
            self._set_per_point(name, center[k])
        self.center_of_voxel = True

    @trace
    @timeit
    def compute_corner_of_voxel(self):
        """corner_* = (voxel * vs) + origin (vapc.py:911-925)."""
        if self.voxelized is False:
            self._voxelize()
            self.drop_columns += ["voxel_x", "voxel_y", "voxel_z"]
        _, corner = self._get_cloud().center_corner(center=False, corner=True)
        for k, name in enumerate(("corner_x", "corner_y", "corner_z")):
            self._set_per_point(name, corner[k])
        self.corner_of_voxel = True

    @trace
    @timeit
    def compute_covariance_matrix(self):
        """The 9 cov_* columns, moved to the front of the frame (vapc.py:929-1040); NaN for
        single-point voxels.  Sets `covariance_matrix_computed` — like the reference it does NOT set
        the `covariance_matrix` flag."""
        if not self.voxelized:
            self._voxelize()
        cloud = self._get_cloud()
        if self.covariance_formula == "reference":
            # the reference's own columns, rounding for rounding (raw-moment formula over Kahan sums in row order): for diffing
            # against its output; ill-conditioned at large coordinates (its errors are reproduced, not corrected)
            cov = cloud.covariance_reference_formula()
        elif self.covariance_formula == "exact":
            cov = cloud.stats(cov=True).cov
        else:
            raise ValueError("covariance_formula must be 'exact' or 'reference'")
        for name, k in zip(_COV_NAMES, _COV9_FROM6):
            self._add_lazy(name, "voxel", cov[k], cloud)
        self._cols = _COV_NAMES + [c for c in self._cols if c not in _COV_NAMES]  # the reference moves them to the front
        self.covariance_matrix_computed = True

    @trace
    @timeit
    def compute_eigenvalues(self):
        """Eigenvalue_1..3 of the per-voxel covariance, descending, NaN for single-point voxels
        (vapc.py:1044-1090; the reference's merge resets the index)."""
        if self.covariance_matrix is False:
            self.compute_covariance_matrix()
            if "covariance_matrix" not in self.compute:
                self.drop_columns += list(_COV_NAMES)
        cloud = self._get_cloud()
        eig = cloud.stats(eig=True).eig
        self._reset_index()
        for k, name in enumerate(("Eigenvalue_1", "Eigenvalue_2", "Eigenvalue_3")):
            self._set_per_point(name, eig[k])
        self.eigenvalues = True

    @trace
    @timeit
    def compute_geometric_features(self):
        """The 8 eigenvalue features plus the Eigenvalue_{1,2,3}_n columns the reference leaves behind
        (vapc.py:1094-1146)."""
        if self.eigenvalues is False:
            self.compute_eigenvalues()
            if "eigenvalues" not in self.compute:
                self.drop_columns += ["Eigenvalue_1", "Eigenvalue_2", "Eigenvalue_3"]
        cloud = self._get_cloud()
        st = cloud.stats(eig_n=True, feat=True)
        for k, name in enumerate(("Eigenvalue_1_n", "Eigenvalue_2_n", "Eigenvalue_3_n")):
            self._set_per_point(name, st.eig_n[k])
        for k, name in enumerate(E.FEATURE_NAMES):
            self._set_per_point(name, st.feat[k])
        self.geometric_features = True

    # ------------------------------------------------------------------------------------
    @trace
    @timeit
    def reduce_to_voxels(self):
        """One row per voxel at the centre / corner / centroid / closest point(s) (vapc.py:1150-1210)."""
        if self.return_at == "center_of_voxel":
            if self.center_of_voxel is False or not self._has("center_x"):
                self.compute_center_of_voxel()
                self.drop_columns += ["center_x", "center_y", "center_z"]
            self.new_column_names.update({"X": "center_x", "Y": "center_y", "Z": "center_z"})
        elif self.return_at == "corner_of_voxel":
            if self.corner_of_voxel is False:
                self.compute_corner_of_voxel()
                self.drop_columns += ["corner_x", "corner_y", "corner_z"]
            self.new_column_names.update({"X": "corner_x", "Y": "corner_y", "Z": "corner_z"})
        elif self.return_at == "center_of_gravity":
            if not self.center_of_gravity:
                self.compute_center_of_gravity()
                self.drop_columns += ["cog_x", "cog_y", "cog_z"]
            self.new_column_names.update({"X": "cog_x", "Y": "cog_y", "Z": "cog_z"})
        elif self.return_at == "closest_to_center_of_gravity":
            if not self.center_of_gravity:
                self.compute_center_of_gravity()
                self.drop_columns += ["cog_x", "cog_y", "cog_z"]
            if not self.closest_to_center_of_gravity:
                self.compute_closest_to_center_of_gravity()
                self.drop_columns += ["min_distance"]
        else:
            print(f"Voxels cannot be reduced to {self.return_at},\n \
                    try 'center_of_gravity', 'center_of_voxel', 'closest_to_center_of_gravity', or 'corner_of_voxel'")
            return
        # From here on the reference works on the frame (vapc.py:1195-1209): [closest: keep the rows with distance == min_distance,]
        # copy the new coordinates into X, Y, Z, drop the helper columns, drop_duplicates(subset=XYZ),
        # groupby(XYZ, as_index=False).median(numeric_only=True).  After the drop every group is ONE row, so the result is: the first
        # row of each distinct (X, Y, Z), sorted by (X, Y, Z), key columns first, every other numeric column as float64.  Here:
        # candidate rows (one per voxel: its first point; or the closest points) -> their X, Y, Z -> three radix sorts on the GPU for
        # the row order -> every output column gathered for those rows only — from its [V] voxel column on the device when it is
        # still pending, from the host column otherwise — into one float64 block that the frame wraps without a copy.
        dropped = set(self.drop_columns)
        missing = [c for c in dropped if not self._has(c)]
        if missing:
            raise KeyError(f"{missing} not found in axis")
        cloud = self._cloud_if_current()
        n = len(self._df)
        dev = cloud.device if cloud is not None else torch.device("cuda", torch.cuda.current_device())
        p2v = _u32(cloud.point2voxel) if cloud is not None else None

        def gather(name, rows_dev, rows_np):
            """Column `name` at the given point rows: a device tensor when the column is pending, else a host array."""
            hit = self._lazy.get(name)
            if hit is not None:
                kind, t, c = hit
                return t[_u32(c.point2voxel)[rows_dev]] if kind == "voxel" else t[rows_dev]
            if cloud is not None and name in ("X", "Y", "Z") and getattr(cloud, "x", None) is not None:
                # the device copy of the coordinates the current cloud was built from (bit-identical to the frame's columns)
                return (cloud.x, cloud.y, cloud.z)["XYZ".index(name)][rows_dev]
            a = np.ascontiguousarray(self._df[name].to_numpy())
            if a.dtype not in (np.float64, np.float32, np.int64, np.int32, np.int16, np.int8, np.uint8):
                a = a.astype(np.float64)  # bool, unsigned 16 / 32 / 64 bit: not gatherable through torch
            with warnings.catch_warnings():
                warnings.simplefilter("ignore")  # read-only numpy views are only read here
                return torch.from_numpy(a).index_select(0, rows_np()).numpy()  # multi-threaded gather

        # ---- candidate rows -------------------------------------------------------------------------------------------
        if self.return_at == "closest_to_center_of_gravity":
            if self.closest_ties == "first" and getattr(self, "_closest_argmin", None) is not None and cloud is not None:
                sel = torch.sort(self._closest_argmin.to(dev)).values
            else:
                allrows = torch.arange(n, device=dev)
                d, m = (gather(c, allrows, lambda: torch.arange(n)) for c in ("distance", "min_distance"))
                d, m = (torch.from_numpy(np.asarray(v, dtype=np.float64)).to(dev) if isinstance(v, np.ndarray) else v for v in (d, m))
                sel = torch.nonzero(d == m).flatten()
                del allrows, d, m
        elif cloud is not None:
            # one row per voxel: its first point (what drop_duplicates(subset=XYZ) keeps, vapc.py:1208)
            sel = torch.sort(_u32(cloud.first_idx)).values
        else:
            sel = torch.arange(n, device=dev)
        sel_np_cache = []

        def sel_np():
            if not sel_np_cache:
                sel_np_cache.append(sel.cpu())
            return sel_np_cache[0]

        xyz = []
        for c in ("X", "Y", "Z"):
            v = gather(self.new_column_names.get(c, c), sel, sel_np)
            xyz.append(torch.from_numpy(np.asarray(v, dtype=np.float64)).to(dev) if isinstance(v, np.ndarray) else v.to(torch.float64))
        rows = E.lexsort_unique_rows(*xyz, device=dev, as_tensor=True)
        final = sel[rows]  # source row of every output row
        final_np_cache = []

        def final_np():
            if not final_np_cache:
                final_np_cache.append(final.cpu())
            return final_np_cache[0]

        # ---- output columns -------------------------------------------------------------------------------------------
        names = [c for c in ["X", "Y", "Z"] + [c for c in self._columns() if c not in ("X", "Y", "Z")] if c not in dropped]
        keep = []
        for c in names:
            src = self.new_column_names.get(c, c)
            if c in ("X", "Y", "Z") or src in self._lazy:
                keep.append(c)
                continue
            dt = self._df[src].dtype
            if pd.api.types.is_bool_dtype(dt) or pd.api.types.is_numeric_dtype(dt):
                keep.append(c)
        block = np.empty((len(keep), final.numel()), dtype=np.float64)
        on_device = [i for i, c in enumerate(keep) if c in ("X", "Y", "Z") or self.new_column_names.get(c, c) in self._lazy]

        def device_column(i):
            c = keep[i]
            return xyz["XYZ".index(c)][rows] if c in ("X", "Y", "Z") else gather(self.new_column_names.get(c, c), final, final_np)

        # the columns that live in the host frame (original attributes) are gathered by a host thread while the device columns come down
        on_host = [i for i in range(len(keep)) if i not in on_device]
        if on_host:
            final_np()  # the row list reaches the host on this thread (its stream), the worker only reads it

        def host_columns():
            for i in on_host:
                block[i] = gather(self.new_column_names.get(keep[i], keep[i]), final, final_np)

        with ThreadPoolExecutor(max_workers=1) as side:
            pending = side.submit(host_columns) if on_host else None
            _download_rows(block, on_device, device_column)
            if pending is not None:
                pending.result()
        self.df = pd.DataFrame(block.T, columns=keep, copy=False)
        self.reduced = True

    @trace
    @timeit
    def filter_attributes(self, filter_attribute: str, min_max_eq: str, filter_value):
        """Row filter == != > >= < <= (long or symbol names, case-insensitive) (vapc.py:1214-1270)."""
        valid_strings = ["equal_to", "==", "not_equal_to", "!=", "greater_than", ">", "greater_than_or_equal_to", ">=",
                         "less_than", "<", "less_than_or_equal_to", "<="]
        op = min_max_eq.lower()
        ops = {"equal_to": "eq", "==": "eq", "not_equal_to": "ne", "!=": "ne", "greater_than": "gt", ">": "gt", "greater_than_or_equal_to": "ge",
               ">=": "ge", "less_than": "lt", "<": "lt", "less_than_or_equal_to": "le", "<=": "le"}
        if not self._has(filter_attribute):
            raise KeyError(filter_attribute)  # a missing column, as in the reference
        if op not in ops:
            raise ValueError(f"Filter invalid, use only one of the following:\n\n{', '.join(valid_strings)}.")
        hit = self._lazy.get(filter_attribute)
        numeric = isinstance(filter_value, (int, float, np.integer, np.floating)) and not isinstance(filter_value, (bool, np.bool_))
        if hit is not None and numeric:
            # the column is still on the device: compare there ([V] voxel values, or [N] point values), compact every pending column
            # to the surviving rows on the device and bring only those back
            kind, t, cloud = hit
            cond = getattr(torch, ops[op])(t, filter_value)
            p2v = _u32(cloud.point2voxel)
            rows = torch.nonzero(cond[p2v] if kind == "voxel" else cond).flatten()
            frame = self._df.iloc[rows.cpu().numpy()]
            new = {}
            for name in self._cols:
                if name in self._lazy:
                    k2, t2, c2 = self._lazy[name]
                    new[name] = self._host(t2[_u32(c2.point2voxel)[rows]] if k2 == "voxel" else t2[rows]).copy()
            order = list(self._cols)
            frame = pd.concat([frame, pd.DataFrame(new, index=frame.index, copy=False)], axis=1)
            self._df = frame if list(frame.columns) == order else frame[order]
        else:
            self._materialise()
            col = self._df[filter_attribute]
            self._df = self._df[getattr(col, ops[op])(filter_value)]
        self._drop_device_state()

    def compute_mode_count(self, attribute: str, percentage: float):
        """Convenience alias named by the north_star: the statistic string "mode_count,<p>" of the
        reference (vapc.py:413-433) for one attribute."""
        saved = self.attributes
        self.attributes = {attribute: [f"mode_count,{percentage}"]}
        self.attributes_up_to_data = True
        try:
            self.compute_requested_statistics_per_attributes()
        finally:
            self.attributes = saved

    @trace
    @timeit
    def compute_clusters(self, cluster_distance, cluster_by=None):
        """cluster_id / cluster_size: rows within `cluster_distance` of each other (in the `cluster_by` columns,
        default the voxel triple) share a cluster (vapc.py:1274-1357).  The reference grows regions with one KD-tree
        query per row in a Python loop; here the rows' connected components come from a union-find over a cell list
        on the GPU (engine.cluster_components), numbered like the reference's sweep numbers them.  Equal to the
        reference whenever no row has more than its `nr_of_neighbours` = 50 rows within the distance."""
        if cluster_by is None:
            cluster_by = ["voxel_x", "voxel_y", "voxel_z"]
        if not self.voxelized:
            self._voxelize()
            self.drop_columns += ["voxel_x", "voxel_y", "voxel_z"]
        missing = [c for c in cluster_by if not self._has(c)]
        if missing:
            raise KeyError(f"{missing} not in index")  # unknown columns, like the reference
        self._materialise(only=tuple(cluster_by))
        cols = self._df[cluster_by]
        if not 1 <= cols.shape[1] <= 3:
            raise NotImplementedError("compute_clusters: one to three cluster_by columns")
        dtype0 = cols.dtypes.iloc[0]
        integer = all(np.issubdtype(t, np.integer) for t in cols.dtypes)
        arrays = [cols.iloc[:, k].to_numpy(np.float64) for k in range(cols.shape[1])]
        arrays += [np.zeros(len(cols))] * (3 - len(arrays))
        label, size, openings = E.cluster_components(*arrays, cluster_distance, integer_rows=integer)
        self.objectCounter = float(openings + 1)
        # np.zeros_like(df[cluster_by[0]]) holds the ids; the sizes come back through a merge (index reset)
        self._assign_host("cluster_id", self._host(label).astype(dtype0))
        self._reset_index()
        self._assign_host("cluster_size", self._host(size).astype(np.result_type(dtype0, np.int64)))
