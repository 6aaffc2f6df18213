"""`DataHandler` — the reference's LAS / LAZ / PLY I/O front end (src/vapc/datahandler.py:12-381) for the
B200 package: same constructor, attributes (`files`, `df`, `las_header`, `attributes`, `voxel_size`) and methods
(`load_las_files`, `save_as_las`, `save_as_ply`), the same DataFrame it hands to `Vapc` (X, Y, Z + every other
LAS dimension, all float64, ExtraBytes dropped; datahandler.py:54-95).

File decoding / encoding is host work next to the hot path.  `laspy` is used when it
is importable (then LAZ works as in the reference); otherwise the native LAS 1.0-1.4 reader / writer in
`vapc_b200.las` handles `.las` and, through the library's own LASzip codec (vapc_laz_decode / vapc_laz_encode), the LAZ
layout of point formats 0 and 2 (+ extra bytes): reading the reference's fixtures, writing `.laz` results.  PLY export needs no third-party
package: the voxel cubes are written as a binary little-endian PLY directly.
"""
from __future__ import annotations

import os
import warnings
from pathlib import Path

import numpy as np
import pandas as pd

from . import las as _las
from .utilities import timeit, trace

try:  # pragma: no cover - not installed in the build image
    import laspy as _laspy
except Exception:  # noqa: BLE001
    _laspy = None

_SKIP_ON_SAVE = ("X", "Y", "Z", "scanner_channel", "overlap")  # datahandler.py:151-152
# 12 triangles of a cube over the corner order of _cube_corner_signs (datahandler.py:338-353)
_CUBE_FACES = np.array([[0, 1, 2], [0, 2, 3], [4, 5, 6], [4, 6, 7], [0, 1, 5], [0, 5, 4], [2, 3, 7], [2, 7, 6], [1, 2, 6], [1, 6, 5],
                        [3, 0, 4], [3, 4, 7]], dtype=np.int32)
_CUBE_SIGNS = np.array([[-1, -1, -1], [1, -1, -1], [1, 1, -1], [-1, 1, -1], [-1, -1, 1], [1, -1, 1], [1, 1, 1], [-1, 1, 1]], dtype=np.float64)


class DataHandler:
    def __init__(self, infiles):
        self.files = infiles if isinstance(infiles, list) else [infiles]
        self.df = None
        self.las_file = None
        self.las_header = None
        self.attributes = None
        self.voxel_size = None

    # ------------------------------------------------------------------------------------------------
    @trace
    @timeit
    def load_las_files(self):
        """Every file -> one float64 DataFrame [X, Y, Z, <other dimensions>], appended to `self.df`
        (datahandler.py:54-95)."""
        frames = []
        if not isinstance(self.files, list):
            self.files = [self.files]
        for path in self.files:
            if not isinstance(path, (str, os.PathLike)):
                # what laspy.open raises for a source that is neither a path nor a stream (tests/integration_tests/test_io.py:134-144)
                raise AttributeError(f"{type(path).__name__!r} object has no attribute 'seekable'")
            path = str(path)
            if _laspy is not None:  # pragma: no cover
                with _laspy.open(path) as fh:
                    data = fh.read()
                    self.las_header = data.header
                    names = [d for d in data.point_format.dimension_names if d not in ("X", "Y", "Z", "ExtraBytes")]
                    cols = {"X": np.asarray(data.x), "Y": np.asarray(data.y), "Z": np.asarray(data.z)}
                    cols.update({d: np.asarray(data[d]) for d in names})
            else:
                header, raw = _las.read_las(path)
                self.las_header = header
                names = [d for d in raw if d not in ("X", "Y", "Z", "x", "y", "z", "ExtraBytes")]
                cols = {"X": raw["x"], "Y": raw["y"], "Z": raw["z"]}
                cols.update({d: raw[d] for d in names})
            self.attributes = names
            frames.append(pd.DataFrame({k: np.asarray(v, dtype=np.float64) for k, v in cols.items()}, columns=["X", "Y", "Z"] + names))
        if self.df is not None:
            frames = [self.df] + frames
        self.df = pd.concat(frames, ignore_index=True)

    # ------------------------------------------------------------------------------------------------
    @trace
    @timeit
    def save_as_las(self, outfile: str, las_point_format=7, las_version="1.4", las_scales=None, las_offset=None):
        """X, Y, Z + every other column (standard dimensions by name, the rest as float32 extra dimensions;
        `scanner_channel` / `overlap` are not written) -> LAS 1.4 (datahandler.py:99-165)."""
        if self.df is None:
            raise AttributeError("DataFrame not found. Please load data before saving.")
        if las_scales is None:
            las_scales = [0.00025, 0.00025, 0.00025]
        xyz = self.df[["X", "Y", "Z"]].to_numpy(np.float64)
        if las_offset is None:
            las_offset = np.min(xyz, axis=0) if len(xyz) else np.zeros(3)
        outfile = str(outfile)
        parent = Path(outfile).parent
        if not os.path.exists(parent):
            os.makedirs(parent)
        if _laspy is not None:  # pragma: no cover
            header = _laspy.LasHeader(point_format=las_point_format, version=las_version)
            header.offsets, header.scales = las_offset, las_scales
            out = _laspy.LasData(header)
            out.x, out.y, out.z = self.df["X"], self.df["Y"], self.df["Z"]
            for name in self.df.columns:
                if name in _SKIP_ON_SAVE:
                    continue
                try:
                    out[name] = self.df[name].astype(np.float32)
                except TypeError:
                    out[name] = self.df[name].astype(np.int16)
                except ValueError:
                    arr = self.df[name].astype(np.float32)
                    out.add_extra_dim(_laspy.ExtraBytesParams(name=name, type=arr.dtype, description=name))
                    out[name] = arr
            self.las_file = out
            out.write(outfile)
            return
        if outfile.lower().endswith(".laz"):
            # native LASzip encoder: LAS 1.4 point format 7 (the reference's default: colour fields always present, zero when the
            # cloud has none — tests/integration_tests/test_io.py:162-180 asserts the format) or 6, + extra bytes
            if str(las_version) != "1.4" or int(las_point_format) not in (6, 7):
                raise NotImplementedError("the native LASzip writer produces LAS 1.4 point formats 6 and 7 (the reference's default is 1.4 / 7)")
            known = set(_las.dimension_names(int(las_point_format)))
            standard, extra = {}, {}
            for name in self.df.columns:
                if name in _SKIP_ON_SAVE:
                    continue
                col = self.df[name].to_numpy()
                if name in known:
                    standard[name] = col
                elif name == "gps_time":
                    extra[name] = np.asarray(col, dtype=np.float64)  # seconds of a GPS week need all 53 bits
                else:
                    extra[name] = np.asarray(col, dtype=np.float64).astype(np.float32)
            _las.write_laz(outfile, xyz[:, 0], xyz[:, 1], xyz[:, 2], standard, extra, scales=las_scales, offsets=las_offset,
                           point_format=int(las_point_format))
            return
        if str(las_version) != "1.4" or int(las_point_format) not in (6, 7, 8):
            raise NotImplementedError("the native writer produces LAS 1.4 point formats 6, 7, 8 (the reference's default is 1.4 / 7)")
        known = set(_las.dimension_names(int(las_point_format)))
        standard, extra = {}, {}
        for name in self.df.columns:
            if name in _SKIP_ON_SAVE:
                continue
            col = self.df[name].to_numpy()
            if name in known:
                standard[name] = col
            else:
                extra[name] = np.asarray(col, dtype=np.float64).astype(np.float32)
        _las.write_las(outfile, xyz[:, 0], xyz[:, 1], xyz[:, 2], standard, extra, scales=las_scales, offsets=las_offset,
                       point_format=int(las_point_format))

    # ------------------------------------------------------------------------------------------------
    def _validate_is_voxelized(self):
        if self.df is None:
            raise AttributeError("DataFrame not found. Please load data before saving.")
        if self.voxel_size is None:
            raise AttributeError("Voxel size not provided. Cannot validate if voxelized.")
        if len(self.df) and len(np.unique(self.df[["X", "Y", "Z"]].to_numpy() % self.voxel_size)) != 1:
            warnings.warn("Coordinates do not look like voxel centres for this voxel size; the cubes are drawn around them anyway.")

    def _calculate_voxel_corners(self, values: np.ndarray) -> np.ndarray:
        """[rows, 3 + k] -> [8 rows, 3 + k]: the 8 cube corners of every row, scalars repeated (datahandler.py:179-230)."""
        half = self.voxel_size / 2.0
        xyz = (values[:, None, :3] + _CUBE_SIGNS[None, :, :] * half).reshape(-1, 3)
        return np.c_[xyz, np.repeat(values[:, 3:], 8, axis=0)]

    def _generate_mesh_data(self):
        values = self.df[["X", "Y", "Z"] + list(self.df.columns[3:])].to_numpy(np.float64)
        corners = self._calculate_voxel_corners(values)
        nvox = len(values)
        faces = np.tile(_CUBE_FACES, (nvox, 1)) + np.repeat(np.arange(nvox, dtype=np.int32) * 8, 12)[:, None]
        return corners, faces.astype(np.int32)

    @trace
    @timeit
    def save_as_ply(self, outfile: str, voxel_size: float, shift_to_center: bool = False):
        """One cube per row of the (voxelised) frame, every other column as a per-vertex float32 property, colours as
        uint8 (16-bit colours rescaled) — binary little-endian PLY (datahandler.py:235-322)."""
        assert voxel_size > 0, "Voxel size must be greater than 0"
        if self.df is None:
            raise AttributeError("DataFrame not found. Please load data before saving.")
        self.voxel_size = voxel_size
        self._validate_is_voxelized()
        self.df.columns = [str(c).replace(" ", "_") for c in self.df.columns]
        if all(c in self.df.columns for c in ("red", "green", "blue")) and len(self.df):
            if max(self.df.red.max(), self.df.green.max(), self.df.blue.max()) > 255:
                for c in ("red", "green", "blue"):
                    self.df[c] = (self.df[c] / 65535.0 * 255).astype(np.uint8)
        verts, faces = self._generate_mesh_data()
        if shift_to_center and len(self.df):
            lo = self.df[["X", "Y", "Z"]].min().to_numpy()
            hi = self.df[["X", "Y", "Z"]].max().to_numpy()
            verts[:, :3] -= lo + (hi - lo) / 2
        columns = ["X", "Y", "Z"] + list(self.df.columns[3:])
        fields = []
        for name in columns:
            fields.append((name.lower() if name in ("X", "Y", "Z") else name, "u1" if name in ("red", "green", "blue") else "<f4"))
        vertex = np.zeros(len(verts), dtype=fields)
        for i, (fname, _) in enumerate(fields):
            with np.errstate(invalid="ignore"):
                vertex[fname] = verts[:, i].astype(vertex.dtype[fname])
        face = np.zeros(len(faces), dtype=[("n", "u1"), ("v", "<i4", (3,))])
        face["n"] = 3
        face["v"] = faces
        ply_types = {"u1": "uchar", "<f4": "float"}
        head = ["ply", "format binary_little_endian 1.0", f"element vertex {len(vertex)}"]
        head += [f"property {ply_types[t]} {n}" for n, t in fields]
        head += [f"element face {len(face)}", "property list uchar int vertex_indices", "end_header"]
        parent = Path(str(outfile)).parent
        if not os.path.exists(parent):
            os.makedirs(parent)
        with open(outfile, "wb") as fh:
            fh.write(("\n".join(head) + "\n").encode("ascii"))
            fh.write(vertex.tobytes())
            fh.write(face.tobytes())
