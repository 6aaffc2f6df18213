"""Host-side I/O next to the hot path (no GPU): the native LAS reader / writer, DataHandler and the PLY export.
Shapes and names follow the reference's tests/integration_tests/test_io.py."""
import json
import os
import struct

import numpy as np
import pandas as pd
import pytest

from vapc_b200 import DataHandler, las

FMT2_ATTRS = ['intensity', 'return_number', 'number_of_returns', 'scan_direction_flag', 'edge_of_flight_line', 'classification', 'synthetic',
              'key_point', 'withheld', 'scan_angle_rank', 'user_data', 'point_source_id', 'red', 'green', 'blue']  # test_io.py:79


def _las12_fmt2(path, n=300, seed=0):
    """A hand-built LAS 1.2 / point format 2 file (the layout of the reference's vapc_in fixture, uncompressed)."""
    rng = np.random.default_rng(seed)
    scale, offset = np.array([0.001, 0.001, 0.001]), np.array([-10.0, 2.0, 0.5])
    X = rng.integers(-5000, 5000, n).astype("<i4")
    Y = rng.integers(0, 3000, n).astype("<i4")
    Z = rng.integers(0, 6000, n).astype("<i4")
    rec = np.zeros(n, dtype=[("X", "<i4"), ("Y", "<i4"), ("Z", "<i4"), ("intensity", "<u2"), ("bits", "u1"), ("cls", "u1"), ("ang", "i1"), ("ud", "u1"),
                             ("psid", "<u2"), ("r", "<u2"), ("g", "<u2"), ("b", "<u2")])
    rec["X"], rec["Y"], rec["Z"] = X, Y, Z
    rec["intensity"] = rng.integers(0, 65535, n)
    rn, nr, sd, ef = rng.integers(1, 6, n), rng.integers(1, 6, n), rng.integers(0, 2, n), rng.integers(0, 2, n)
    rec["bits"] = rn | (nr << 3) | (sd << 6) | (ef << 7)
    cl, syn, kp, wh = rng.integers(0, 32, n), rng.integers(0, 2, n), rng.integers(0, 2, n), rng.integers(0, 2, n)
    rec["cls"] = cl | (syn << 5) | (kp << 6) | (wh << 7)
    rec["ang"] = rng.integers(-90, 90, n)
    rec["ud"] = rng.integers(0, 255, n)
    rec["psid"] = rng.integers(0, 65535, n)
    rec["r"], rec["g"], rec["b"] = (rng.integers(0, 65535, n) for _ in range(3))
    header = bytearray(227)
    header[0:4] = b"LASF"
    header[24], header[25] = 1, 2
    header[58:58 + 6] = b"pytest"
    struct.pack_into("<HII", header, 94, 227, 227, 0)
    struct.pack_into("<BH", header, 104, 2, rec.dtype.itemsize)
    struct.pack_into("<I", header, 107, n)
    x, y, z = X * scale[0] + offset[0], Y * scale[1] + offset[1], Z * scale[2] + offset[2]
    struct.pack_into("<12d", header, 131, *scale, *offset, x.max(), x.min(), y.max(), y.min(), z.max(), z.min())
    with open(path, "wb") as fh:
        fh.write(bytes(header))
        fh.write(rec.tobytes())
    truth = dict(X=x, Y=y, Z=z, intensity=rec["intensity"], return_number=rn, number_of_returns=nr, scan_direction_flag=sd, edge_of_flight_line=ef,
                 classification=cl, synthetic=syn, key_point=kp, withheld=wh, scan_angle_rank=rec["ang"], user_data=rec["ud"],
                 point_source_id=rec["psid"], red=rec["r"], green=rec["g"], blue=rec["b"])
    return truth


def test_read_las12_format2_matches_laspy_conventions(tmp_path):
    path = str(tmp_path / "in.las")
    truth = _las12_fmt2(path)
    dh = DataHandler(path)
    dh.load_las_files()
    assert dh.attributes == FMT2_ATTRS
    assert list(dh.df.columns) == ["X", "Y", "Z"] + FMT2_ATTRS
    assert dh.df.shape == (300, 18) and set(dh.df.dtypes) == {np.dtype("float64")}
    for name, ref in truth.items():
        np.testing.assert_array_equal(dh.df[name].to_numpy(), np.asarray(ref, dtype=np.float64), err_msg=name)  # x = X * scale + offset, two roundings
    assert dh.las_header.version == (1, 2) and dh.las_header.point_format == 2 and dh.las_header.point_count == 300
    np.testing.assert_allclose(dh.las_header.mins, [truth[c].min() for c in "XYZ"])


def test_read_las_ints_are_the_stored_coordinates(tmp_path):
    """las.read_las_ints: the int32 X, Y, Z as stored + the header transform reproduce the DataHandler's float64 columns bit for
    bit with one multiply and one add (the contract of the 12 B/point input path)."""
    path = str(tmp_path / "in.las")
    truth = _las12_fmt2(path)
    header, X, Y, Z = las.read_las_ints(path)
    assert X.dtype == np.int32 and X.flags.c_contiguous and len(X) == 300
    for q, k, name in ((X, 0, "X"), (Y, 1, "Y"), (Z, 2, "Z")):
        np.testing.assert_array_equal(q.astype(np.float64) * header.scales[k] + header.offsets[k], truth[name])


def test_load_several_files_and_append(tmp_path):
    a, b = str(tmp_path / "a.las"), str(tmp_path / "b.las")
    _las12_fmt2(a, 120, 1)
    _las12_fmt2(b, 80, 2)
    dh = DataHandler([a, b])
    dh.load_las_files()
    assert dh.df.shape == (200, 18) and list(dh.df.index) == list(range(200))
    dh.load_las_files()  # loading again appends (datahandler.py:92-95)
    assert dh.df.shape == (400, 18)
    with pytest.raises(FileNotFoundError):
        DataHandler(str(tmp_path / "missing.las")).load_las_files()


def test_save_as_las_roundtrip(tmp_path):
    rng = np.random.default_rng(3)
    n = 5000
    df = pd.DataFrame({"X": rng.uniform(476850, 476900, n), "Y": rng.uniform(5429100, 5429150, n), "Z": rng.uniform(300, 320, n),
                       "intensity": rng.integers(0, 60000, n).astype(float), "classification": rng.integers(0, 20, n).astype(float),
                       "return_number": rng.integers(1, 7, n).astype(float), "gps_time": rng.uniform(0, 1e5, n),
                       "point_count": rng.integers(1, 50, n).astype(float), "cov_xx": rng.normal(0, 1e-3, n), "scanner_channel": np.ones(n),
                       "overlap": np.zeros(n)})
    dh = DataHandler("unused")
    with pytest.raises(AttributeError):
        dh.save_as_las(str(tmp_path / "x.las"))
    dh.df = df
    out = str(tmp_path / "sub" / "out.las")  # parent directory is created
    dh.save_as_las(out)
    back = DataHandler(out)
    back.load_las_files()
    h = back.las_header
    assert h.version == (1, 4) and h.point_format == 7 and h.point_count == n
    np.testing.assert_allclose(h.scales, [0.00025] * 3)
    np.testing.assert_allclose(h.offsets, df[["X", "Y", "Z"]].min().to_numpy())
    for c in "XYZ":
        assert np.abs(back.df[c].to_numpy() - df[c].to_numpy()).max() <= 0.000125 + 1e-9
    for c in ("intensity", "classification", "return_number"):
        np.testing.assert_array_equal(back.df[c].to_numpy(), df[c].to_numpy())
    np.testing.assert_array_equal(back.df["gps_time"].to_numpy(), df["gps_time"].to_numpy())
    for c in ("point_count", "cov_xx"):  # extra dimensions are float32 (datahandler.py:157)
        np.testing.assert_array_equal(back.df[c].to_numpy(), df[c].to_numpy().astype(np.float32).astype(np.float64))
    assert back.df["scanner_channel"].eq(0).all() and back.df["overlap"].eq(0).all()  # not written (datahandler.py:152)
    # .laz: the library's own LASzip encoder and decoder (LAS 1.4 point format 7 like the reference's default — colour fields present,
    # all zero here — + extra bytes, layered; point format 6 on request)
    laz = str(tmp_path / "out.laz")
    dh.save_as_las(laz)
    assert os.path.getsize(laz) < os.path.getsize(out)
    hz, _ = las.read_las(laz)
    assert hz.compressed and hz.version == (1, 4) and hz.point_format == 7 and hz.point_count == n
    assert hz.laszip["compressor"] == 3 and hz.laszip["items"][:2] == [(10, 30, 3), (11, 6, 3)] and hz.laszip["items"][-1][0] == 14
    laz6 = str(tmp_path / "out6.laz")
    dh.save_as_las(laz6, las_point_format=6)
    h6, c6 = las.read_las(laz6)
    assert h6.compressed and h6.point_format == 6 and h6.laszip["items"][0] == (10, 30, 3) and h6.laszip["items"][1][0] == 14 and "red" not in c6
    with pytest.raises(NotImplementedError):
        dh.save_as_las(str(tmp_path / "out8.laz"), las_point_format=8)
    again = DataHandler(laz)
    again.load_las_files()
    for c in "XYZ":
        assert np.abs(again.df[c].to_numpy() - df[c].to_numpy()).max() <= 0.000125 + 1e-9
    for c in ("intensity", "classification", "return_number"):
        np.testing.assert_array_equal(again.df[c].to_numpy(), df[c].to_numpy())
    np.testing.assert_array_equal(again.df["gps_time"].to_numpy(), df["gps_time"].to_numpy())
    for c in ("point_count", "cov_xx"):
        np.testing.assert_array_equal(again.df[c].to_numpy(), df[c].to_numpy().astype(np.float32).astype(np.float64))
    plain = str(tmp_path / "plain.laz")
    os.rename(out, plain)  # an uncompressed file under a .laz name is read as what it is
    assert las.read_las(plain)[0].point_count == n


def test_extra_bytes_vlr_types_and_empty_cloud(tmp_path):
    n = 17
    x = np.linspace(0, 1, n)
    extra = {"a_f32": np.arange(n, dtype=np.float32), "b_i16": np.arange(n, dtype=np.int16) - 5, "c_f64": np.arange(n) * 0.1, "d_u8": np.arange(n, dtype=np.uint8)}
    p = str(tmp_path / "e.las")
    las.write_las(p, x, x * 2, x * 3, {"user_data": np.arange(n)}, extra)
    h, cols = las.read_las(p)
    assert [name for name, _ in h.extra_dims] == list(extra)
    for k, v in extra.items():
        np.testing.assert_array_equal(cols[k], v)
        assert cols[k].dtype == v.dtype
    p0 = str(tmp_path / "empty.las")
    las.write_las(p0, [], [], [], {}, {})
    h0, c0 = las.read_las(p0)
    assert h0.point_count == 0 and len(c0["x"]) == 0
    with pytest.raises(ValueError):
        las.read_header(b"not a las file" + bytes(400))


def test_save_as_ply_cubes(tmp_path):
    vs = 0.5
    df = pd.DataFrame({"X": [0.25, 0.75, 0.25], "Y": [0.25, 0.25, 0.75], "Z": [0.25, 0.25, 0.25], "point count": [3.0, 7.0, 1.0],
                       "red": [65535.0, 0.0, 32768.0], "green": [0.0, 65535.0, 0.0], "blue": [0.0, 0.0, 65535.0]})
    dh = DataHandler("unused")
    dh.df = df
    out = str(tmp_path / "v.ply")
    dh.save_as_ply(out, vs)
    raw = open(out, "rb").read()
    head, body = raw.split(b"end_header\n", 1)
    lines = head.decode().splitlines()
    assert lines[:3] == ["ply", "format binary_little_endian 1.0", "element vertex 24"]
    props = [ln.split()[1:] for ln in lines if ln.startswith("property") and "list" not in ln]
    assert props == [["float", "x"], ["float", "y"], ["float", "z"], ["float", "point_count"], ["uchar", "red"], ["uchar", "green"], ["uchar", "blue"]]
    assert "element face 36" in lines
    vdt = np.dtype([("x", "<f4"), ("y", "<f4"), ("z", "<f4"), ("point_count", "<f4"), ("red", "u1"), ("green", "u1"), ("blue", "u1")])
    v = np.frombuffer(body, dtype=vdt, count=24)
    f = np.frombuffer(body, dtype=np.dtype([("n", "u1"), ("v", "<i4", (3,))]), count=36, offset=24 * vdt.itemsize)
    assert len(body) == 24 * vdt.itemsize + 36 * 13
    np.testing.assert_allclose(sorted(set(v["x"][:8])), [0.0, 0.5])  # first cube spans [0, 0.5]^3
    assert set(v["point_count"][:8]) == {3.0} and set(v["red"][:8]) == {255} and set(v["red"][16:]) == {127}
    assert f["n"].tolist() == [3] * 36 and f["v"][:12].max() == 7 and f["v"][12:24].min() == 8
    # every cube face is two triangles over 4 coplanar corners
    tri = v[["x", "y", "z"]][f["v"][:2].ravel()]
    assert len({t[2] for t in tri.tolist()}) == 1


def test_command_line_templates_schema():
    root = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "command_line_templates")
    names = sorted(os.listdir(root))
    assert names == ["compute_attributes_config.json", "compute_statistics_config.json", "filter_and_compute_attributes.json",
                     "filter_by_attributes_config.json", "mask_config.json", "subsample_config.json"]
    tools = set()
    for nme in names:
        cfg = json.load(open(os.path.join(root, nme)))
        assert set(cfg) == {"infile", "outdir", "voxel_size", "vapc_command", "tile", "reduce_to", "save_as"}  # main.py:228-236
        assert set(cfg["vapc_command"]) == {"tool", "args"}
        tools.add(cfg["vapc_command"]["tool"])
    assert tools == {"compute", "statistics", "filter_and_compute", "filter", "mask", "subsample"}  # vapc_tools.py:376-411


def test_save_as_laz_with_colours_uses_point_format_7(tmp_path):
    """Frames with red / green / blue are written as LAS 1.4 point format 7 + extra bytes (POINT14 v3 + RGB14 v3 + BYTE14 v3):
    the reference's own output format (datahandler.py:99-165)."""
    rng = np.random.default_rng(2)
    n = 5000
    df = pd.DataFrame({"X": np.round(rng.random(n) * 30, 3), "Y": np.round(rng.random(n) * 30, 3), "Z": np.round(rng.random(n) * 5, 3),
                       "intensity": rng.integers(0, 4000, n).astype(float), "classification": rng.integers(0, 20, n).astype(float),
                       "red": rng.integers(0, 65536, n).astype(float), "green": rng.integers(0, 65536, n).astype(float),
                       "blue": rng.integers(0, 65536, n).astype(float), "gps_time": 3e5 + np.arange(n) * 1e-5, "point_count": rng.integers(1, 9, n).astype(float)})
    dh = DataHandler("")
    dh.df = df.copy()
    path = str(tmp_path / "rgb.laz")
    dh.save_as_las(path)
    h, _ = las.read_las(path)
    assert h.compressed and h.version == (1, 4) and h.point_format == 7 and h.laszip["items"][:2] == [(10, 30, 3), (11, 6, 3)] and h.laszip["items"][2][0] == 14
    back = DataHandler(path)
    back.load_las_files()
    for c in ("intensity", "classification", "red", "green", "blue", "gps_time", "point_count"):
        np.testing.assert_array_equal(back.df[c].to_numpy(), df[c].to_numpy(), err_msg=c)
    for c in "XYZ":
        assert np.abs(back.df[c].to_numpy() - df[c].to_numpy()).max() <= 0.000125 + 1e-9


def test_las_to_laz_conversion_of_the_command_line(tmp_path):
    """What main.do_vapc_on_files does for save_as=".laz" (main.py:154-168 of the reference): the tool's LAS 1.4 / format 7 result is
    loaded again and saved as .laz: LAS 1.4 point format 7 again (colour dimensions all zero: the cloud had none) + extra bytes;
    every column survives."""
    rng = np.random.default_rng(4)
    n = 3000
    df = pd.DataFrame({"X": np.round(rng.random(n) * 30, 3), "Y": np.round(rng.random(n) * 30, 3), "Z": np.round(rng.random(n) * 5, 3),
                       "intensity": rng.integers(0, 4000, n).astype(float), "point_count": rng.integers(1, 9, n).astype(float),
                       "cog_x": rng.random(n), "cog_y": rng.random(n), "cog_z": rng.random(n)})
    first = DataHandler("")
    first.df = df.copy()
    las_path = str(tmp_path / "compute.las")
    first.save_as_las(las_path)
    again = DataHandler(las_path)
    again.load_las_files()
    assert {"red", "green", "blue"} <= set(again.df.columns)  # format 7 always carries them
    laz_path = str(tmp_path / "compute.laz")
    again.save_as_las(laz_path)
    h, _ = las.read_las(laz_path)
    assert h.compressed and h.version == (1, 4) and h.point_format == 7 and os.path.getsize(laz_path) < os.path.getsize(las_path)
    res = DataHandler(laz_path)
    res.load_las_files()
    assert len(res.df) == n and {"point_count", "cog_x", "cog_y", "cog_z", "intensity", "red", "green", "blue"} <= set(res.df.columns)
    assert not res.df[["red", "green", "blue"]].to_numpy().any()
    for c in ("point_count", "cog_x", "cog_y", "cog_z", "intensity"):
        np.testing.assert_array_equal(res.df[c].to_numpy(), again.df[c].to_numpy(), err_msg=c)
    for c in "XYZ":
        assert np.abs(res.df[c].to_numpy() - df[c].to_numpy()).max() <= 2 * 0.000125 + 1e-9


def test_writer_refuses_coordinates_that_do_not_fit_int32(tmp_path):
    """laspy raises OverflowError where the int32 record cannot hold (c - offset) / scale; so does the native writer (no silent wrap)."""
    from vapc_b200 import las

    x = np.array([0.0, 600_000.0])  # 600 km at the default 0.25 mm scale: 2.4e9 units
    with pytest.raises(OverflowError):
        las.write_las(str(tmp_path / "far.las"), x, x * 0, x * 0, {}, {})
    with pytest.raises(OverflowError):
        las.write_las(str(tmp_path / "nan.las"), np.array([0.0, np.nan]), np.zeros(2), np.zeros(2), {}, {})


def test_extra_bytes_description_in_an_evlr(tmp_path):
    """LAS 1.4 allows the extra-bytes description (LASF_Spec / 4) to live in an EXTENDED VLR behind the points: the reader finds it there."""
    import struct

    from vapc_b200 import las

    rng = np.random.default_rng(3)
    n = 50
    x, y, z = (np.round(rng.uniform(0, 10, n), 3) for _ in range(3))
    amp = rng.normal(size=n).astype(np.float32)
    path = str(tmp_path / "vlr.las")
    las.write_las(path, x, y, z, {}, {"amplitude": amp})
    buf = bytearray(open(path, "rb").read())
    header_size, offset_to_points, n_vlr = struct.unpack_from("<HII", buf, 94)
    assert n_vlr == 1
    vlr = bytes(buf[header_size:offset_to_points])  # the 54-byte VLR header + the 192-byte description
    body = vlr[54:]
    evlr = struct.pack("<H16sHQ32s", 0, b"LASF_Spec", 4, len(body), b"extra bytes") + body
    points = bytes(buf[offset_to_points:])
    head = bytearray(buf[:header_size])
    struct.pack_into("<HII", head, 94, header_size, header_size, 0)  # no VLRs: the points follow the header
    struct.pack_into("<QI", head, 235, header_size + len(points), 1)  # one EVLR behind the points
    moved = str(tmp_path / "evlr.las")
    open(moved, "wb").write(bytes(head) + points + evlr)
    h, rec = las.read_las(moved)
    assert ("amplitude", "<f4") in [(name, np.dtype(dt).str) for name, dt in h.extra_dims]
    np.testing.assert_array_equal(rec["amplitude"], amp)
