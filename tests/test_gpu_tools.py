"""Tool layer + JSON command line on the GPU (reference: tests/integration_tests/test_main.py, test_templates.py):
files in, files out, checked against the CPU oracle on the same points."""
import glob
import json
import os

import numpy as np
import pandas as pd
import pytest

from oracle import vapc_oracle as O

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def V():
    import vapc_b200

    return vapc_b200


def _write_cloud(V, path, n=40000, seed=5, extent=(12.0, 9.0, 3.0), with_rgb=True):
    rng = np.random.default_rng(seed)
    # multiples of the default LAS scale (2.5e-4) so that the file round trip is exact
    q = [rng.integers(0, int(e / 0.00025), n) for e in extent]
    df = pd.DataFrame({"X": q[0] * 0.00025, "Y": q[1] * 0.00025, "Z": q[2] * 0.00025, "intensity": rng.integers(0, 5000, n).astype(float),
                       "classification": rng.integers(1, 9, n).astype(float)})
    if with_rgb:
        for c in ("red", "green", "blue"):
            df[c] = rng.integers(0, 256, n).astype(float)
    dh = V.DataHandler("unused")
    dh.df = df.copy()
    dh.save_as_las(path, las_offset=[0.0, 0.0, 0.0])
    return df


def _read(V, path):
    dh = V.DataHandler(path)
    dh.load_las_files()
    return dh.df


def test_use_tool_compute_center_of_voxel(V, tmp_path):
    src, out = str(tmp_path / "in.las"), str(tmp_path / "out.las")
    df = _write_cloud(V, src)
    vs = 0.5
    V.use_tool("compute", src, out, voxel_size=vs, args={"compute": ["point_count", "point_density", "center_of_gravity", "covariance_matrix"]},
               reduce_to="center_of_voxel")
    res = _read(V, out)
    x, y, z = (df[c].to_numpy() for c in "XYZ")
    g = O.group_by_voxel(*O.voxelize(x, y, z, [0, 0, 0], vs))
    assert len(res) == g.nvoxels
    center = O.center_of_voxel(g, [0, 0, 0], vs)
    order = np.lexsort((center[:, 2], center[:, 1], center[:, 0]))  # reduce_to_voxels sorts by X, Y, Z (vapc.py:1208-1209)
    np.testing.assert_allclose(res[["X", "Y", "Z"]].to_numpy(), center[order], atol=0.000126)
    np.testing.assert_array_equal(res["point_count"].to_numpy(), g.counts[order])
    np.testing.assert_allclose(res["point_density"].to_numpy(), (g.counts / vs ** 3)[order], rtol=1e-6)
    cog = O.center_of_gravity(x, y, z, g)[order]
    np.testing.assert_allclose(res[["cog_x", "cog_y", "cog_z"]].to_numpy(), cog, rtol=2e-7)  # float32 extra dimensions
    for c in ("cov_xx", "cov_yz", "cov_zz"):
        assert c in res.columns


def test_use_tool_filter_statistics_subsample(V, tmp_path):
    src = str(tmp_path / "in.las")
    df = _write_cloud(V, src)
    x, y, z = (df[c].to_numpy() for c in "XYZ")
    vs = 1.0
    g = O.group_by_voxel(*O.voxelize(x, y, z, [0, 0, 0], vs))
    # filter: keeps the points of voxels with more than 100 points, then one row per voxel
    out = str(tmp_path / "filter.las")
    V.use_tool("filter", src, out, voxel_size=vs, args={"filters": {"point_count": {"greater_than": 100}}}, reduce_to="center_of_voxel")
    res = _read(V, out)
    assert len(res) == int((g.counts > 100).sum()) and (res["point_count"] > 100).all()
    # an operator the tool does not know makes it return False, the frame is then not a frame (vapc_tools.py:218-225)
    vp, dh = V.vapc_tools.initiate_vapc(src, vs)
    assert V.vapc_tools.filter_by_attributes(vp, {"point_count": {"not_equal_to": 3}}) is False
    # statistics: mode / mode_count / mean of an integer attribute
    out = str(tmp_path / "stats.las")
    V.use_tool("statistics", src, out, voxel_size=vs, args={"statistics": {"classification": ["mode", "mode_count,0.1", "mean"]}}, reduce_to="center_of_voxel")
    res = _read(V, out)
    cls = df["classification"].to_numpy()
    mode, mc = O.attribute_statistic(cls, g, "mode"), O.attribute_statistic(cls, g, "mode_count,0.1")
    center = O.center_of_voxel(g, [0, 0, 0], vs)
    order = np.lexsort((center[:, 2], center[:, 1], center[:, 0]))
    np.testing.assert_array_equal(res["classification_mode"].to_numpy(), mode[order])
    np.testing.assert_array_equal(res["classification_mode_count,0.1"].to_numpy(), mc[order])
    mean = np.bincount(g.point2voxel, weights=cls) / g.counts
    np.testing.assert_allclose(res["classification_mean"].to_numpy(), mean[order], rtol=2e-7)
    # subsample: one original point per voxel
    out = str(tmp_path / "sub.las")
    V.use_tool("subsample", src, out, voxel_size=vs, args={"sub_sample_method": "closest_to_center_of_gravity"}, reduce_to=False)
    res = _read(V, out)
    assert g.nvoxels <= len(res) <= g.nvoxels + 5  # exact ties keep more than one point (vapc.py:1195)
    pts = {tuple(np.round(p / 0.00025).astype(np.int64)) for p in df[["X", "Y", "Z"]].to_numpy()}
    assert all(tuple(np.round(p / 0.00025).astype(np.int64)) in pts for p in res[["X", "Y", "Z"]].to_numpy())
    with pytest.raises(ValueError):
        V.use_tool("no_such_tool", src, out, voxel_size=vs, args={}, reduce_to=False)
    assert V.use_tool("subsample", src, str(tmp_path / "x.xyz"), voxel_size=vs, args={"sub_sample_method": "center_of_voxel"}, reduce_to=False) == "Unknown output format"


def test_use_tool_mask(V, tmp_path):
    src, msk = str(tmp_path / "in.las"), str(tmp_path / "mask.las")
    df = _write_cloud(V, src, n=30000, seed=7)
    mdf = _write_cloud(V, msk, n=300, seed=8, extent=(4.0, 4.0, 2.0), with_rgb=False)
    vs = 0.5
    vx, vy, vz = O.voxelize(*(df[c].to_numpy() for c in "XYZ"), [0, 0, 0], vs)
    mv = set(zip(*O.voxelize(*(mdf[c].to_numpy() for c in "XYZ"), [0, 0, 0], vs)))
    for buf, mode in ((0, "in"), (1, "in"), (0, "out")):
        cells = set()
        for (a, b, c) in mv:
            for i in range(-buf, buf + 1):
                for j in range(-buf, buf + 1):
                    for k in range(-buf, buf + 1):
                        cells.add((a + i, b + j, c + k))
        inside = np.array([(a, b, c) in cells for a, b, c in zip(vx, vy, vz)])
        keep = inside if mode == "in" else ~inside
        out = str(tmp_path / f"m_{buf}_{mode}.las")
        V.use_tool("mask", src, out, voxel_size=vs, args={"maskfile": msk, "segment_in_or_out": mode, "buffer_size": buf}, reduce_to=False)
        res = _read(V, out)
        assert len(res) == int(keep.sum())
        np.testing.assert_allclose(res[["X", "Y", "Z"]].to_numpy(), df[["X", "Y", "Z"]].to_numpy()[keep], atol=0.000126)  # original order kept


@pytest.mark.parametrize("save_as", [".las", ".ply", ".laz"])
def test_do_vapc_on_files_and_main(V, tmp_path, save_as, capsys):
    src = str(tmp_path / "in.las")
    df = _write_cloud(V, src, n=20000)
    outdir = str(tmp_path / "out")
    cfg = {"infile": src, "outdir": outdir, "voxel_size": 1.0, "vapc_command": {"tool": "compute", "args": {"compute": ["point_count", "center_of_gravity"]}},
           "tile": 20, "reduce_to": "center_of_voxel", "save_as": save_as}
    cfg_path = str(tmp_path / "cfg.json")
    json.dump(cfg, open(cfg_path, "w"))
    written = V.main([cfg_path])
    assert written.endswith(save_as)  # .laz: compressed by the library's own LASzip encoder (LAS 1.2 + extra bytes)
    assert os.path.exists(written) and os.path.basename(written).startswith("compute_")
    docs = glob.glob(os.path.join(outdir, "compute_*_cfg.json"))
    assert len(docs) == 1
    doc = json.load(open(docs[0]))
    assert doc["tool"] == "compute" and doc["voxel_size"] == 1.0 and doc["reduce_to"] == "center_of_voxel" and doc["tile"] == 20 and "vapc_version" in doc
    g = O.group_by_voxel(*O.voxelize(*(df[c].to_numpy() for c in "XYZ"), [0, 0, 0], 1.0))
    if written.endswith((".las", ".laz")):
        if save_as == ".laz":
            assert V.las.read_las(written)[0].compressed and not os.path.exists(written.replace(".laz", ".las"))
        res = _read(V, written)
        assert len(res) == g.nvoxels and {"point_count", "cog_x", "cog_y", "cog_z"} <= set(res.columns)
    else:
        head = open(written, "rb").read(2000).split(b"end_header")[0].decode()
        assert f"element vertex {8 * g.nvoxels}" in head and "property float point_count" in head and "property float cog_z" in head


def test_do_vapc_on_files_errors_and_file_list(V, tmp_path):
    a, b = str(tmp_path / "a.las"), str(tmp_path / "b.las")
    da = _write_cloud(V, a, n=5000, seed=1)
    db = _write_cloud(V, b, n=7000, seed=2)
    cmd = {"tool": "compute", "args": {"compute": ["point_density", "point_count", "percentage_occupied"]}}
    with pytest.raises(FileNotFoundError):
        V.do_vapc_on_files(file=str(tmp_path / "missing.las"), out_dir=str(tmp_path / "o1"), voxel_size=0.5, vapc_command=dict(cmd))
    with pytest.raises(TypeError):
        V.do_vapc_on_files(file=a, out_dir=str(tmp_path / "o1"), voxel_size=0.5, vapc_command=dict(cmd), tilesize=20)
    with pytest.raises(ValueError):
        V.do_vapc_on_files(file=a, out_dir=str(tmp_path / "o1"), voxel_size=0.5, vapc_command={"tool": "invalid_tool", "args": {"compute": ["point_density"]}})
    with pytest.raises(ValueError):
        V.do_vapc_on_files(file=a, out_dir=str(tmp_path / "o1"), voxel_size=0.5, vapc_command={"tool": "compute", "args": {"compute": ["median"]}})
    out = V.do_vapc_on_files(file=[a, b], out_dir=str(tmp_path / "o2"), voxel_size=1.0, vapc_command=dict(cmd), save_as=".las")
    both = pd.concat([da, db], ignore_index=True)
    g = O.group_by_voxel(*O.voxelize(*(both[c].to_numpy() for c in "XYZ"), [0, 0, 0], 1.0))
    res = _read(V, out)
    assert len(res) >= g.nvoxels and "point_count" in res.columns  # default reduction: closest point(s) per voxel
