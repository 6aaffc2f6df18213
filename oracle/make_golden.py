"""Generate tests/golden/*.npz by running the UNMODIFIED reference (3dgeo-heidelberg/vapc).

Run in the build container only (needs /root/reference):

    python oracle/make_golden.py

TEST INFRASTRUCTURE.  The fixtures are committed; the GPU box never runs this script.
Each .npz holds the inputs (`in_*`) and the reference's outputs, column by column, in
original point order (`pt_*`), the four `reduce_to_voxels` results (`red_<mode>_*`), the
attribute statistics (`stat_*`), mask selections (`mask_*`) and — for the two-epoch cases —
the merged / exported frames (`bi_*`).  Reference versions are recorded in `meta_versions`.
"""
import os
import sys
import warnings

import numpy as np
import pandas as pd

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
from oracle.ref_harness import import_reference  # noqa: E402

warnings.filterwarnings("ignore")
vapc = import_reference()
OUT = os.path.join(os.path.dirname(HERE), "tests", "golden")

FULL = ["point_count", "point_density", "center_of_gravity", "std_of_cog", "covariance_matrix",
        "eigenvalues", "geometric_features", "distance_to_center_of_gravity",
        "closest_to_center_of_gravity", "center_of_voxel", "corner_of_voxel"]
MODES = ["center_of_voxel", "corner_of_voxel", "center_of_gravity", "closest_to_center_of_gravity"]


def new_vapc(df, vs, origin=None, **kw):
    v = vapc.Vapc(vs, origin=list(origin) if origin is not None else None, **kw)
    v.df = df.copy()
    v.original_attributes = v.df.columns.tolist()
    return v


def cols_to(store, prefix, df):
    for c in df.columns:
        store[f"{prefix}{c}"] = df[c].to_numpy()


def run_cloud(store, df, vs, origin=None, stats=None, mask_df=None, buffer_size=1):
    for c in df.columns:
        store[f"in_{c}"] = df[c].to_numpy()
    store["in_voxel_size"] = np.float64(vs)
    store["in_origin"] = np.asarray(origin if origin is not None else [0, 0, 0], np.float64)
    # per-point attributes, full set
    v = new_vapc(df, vs, origin)
    v.compute = list(FULL)
    v.compute_requested_attributes()
    cols_to(store, "pt_", v.df.reset_index(drop=True))
    store["pt__columns"] = np.array(v.df.columns.tolist())
    v2 = new_vapc(df, vs, origin)
    v2.compute_percentage_occupied()
    store["percentage_occupied"] = np.float64(v2.percentage_occupied)
    # reductions
    for mode in MODES:
        r = new_vapc(df, vs, origin, return_at=mode)
        r.reduce_to_voxels()
        cols_to(store, f"red_{mode}_", r.df)
        store[f"red_{mode}__columns"] = np.array(r.df.columns.tolist())
    # per-attribute statistics
    if stats:
        s = new_vapc(df, vs, origin, attributes=stats)
        s.compute_requested_statistics_per_attributes()
        for c in s.df.columns:
            if c not in df.columns:
                store[f"stat_{c}"] = s.df[c].to_numpy()
    # mask
    if mask_df is not None:
        for c in ("X", "Y", "Z"):
            store[f"in_mask_{c}"] = mask_df[c].to_numpy()
        for mode in ("in", "out"):
            p = new_vapc(df.assign(_row=np.arange(len(df), dtype=np.float64)), vs, origin)
            m = new_vapc(mask_df, vs, origin)
            p.select_by_mask(m, segment_in_or_out=mode)
            store[f"mask_{mode}_rows"] = p.df["_row"].to_numpy().astype(np.int64)
        m = new_vapc(mask_df, vs, origin)
        buf = m.compute_voxel_buffer(buffer_size=buffer_size)
        store["mask_buffer_voxels"] = buf[["voxel_x", "voxel_y", "voxel_z"]].to_numpy().astype(np.int64)
        store["mask_buffer_xyz"] = buf[["X", "Y", "Z"]].to_numpy()
        store["in_mask_buffer_size"] = np.int64(buffer_size)
        p = new_vapc(df.assign(_row=np.arange(len(df), dtype=np.float64)), vs, origin)
        m.voxel_index = False
        m.df = m.buffer_df
        p.select_by_mask(m, segment_in_or_out="in")
        store["mask_buffer_in_rows"] = p.df["_row"].to_numpy().astype(np.int64)


def run_bitemporal(store, df1, df2, vs, alpha=0.005):
    for k, d in (("1", df1), ("2", df2)):
        for c in ("X", "Y", "Z"):
            store[f"in_e{k}_{c}"] = d[c].to_numpy()
    store["in_bi_voxel_size"] = np.float64(vs)
    store["in_bi_alpha"] = np.float64(alpha)
    a = new_vapc(df1[["X", "Y", "Z"]], vs)
    b = new_vapc(df2[["X", "Y", "Z"]], vs)
    bi = vapc.BiTemporalVapc([a, b])
    bi.prepare_data_for_mahalanobis_distance()
    bi.merge_vapcs_with_same_voxel_index()
    bi.compute_distance()
    m0 = bi.merged_df.copy()
    # d1/d2 are internal temporaries of bi_vapc.py:116-117; recompute them with the same
    # numpy expressions from the reference's merged frame so the kernels can be checked.
    x1 = m0[["cog_x_x", "cog_y_x", "cog_z_x"]].to_numpy()
    x2 = m0[["cog_x_y", "cog_y_y", "cog_z_y"]].to_numpy()
    cx = m0[sorted(c for c in m0.columns if c.startswith("cov_") and c.endswith("_x"))].to_numpy().reshape(-1, 3, 3).copy()
    cy = m0[sorted(c for c in m0.columns if c.startswith("cov_") and c.endswith("_y"))].to_numpy().reshape(-1, 3, 3).copy()
    cx += np.eye(3) * 1e-10
    cy += np.eye(3) * 1e-10
    diff = x1 - x2
    with np.errstate(all="ignore"):
        try:
            store["bi_d1"] = np.einsum("ij,ijk,ik->i", diff, np.linalg.inv(cy), diff)
            store["bi_d2"] = np.einsum("ij,ijk,ik->i", diff, np.linalg.inv(cx), diff)
        except np.linalg.LinAlgError:
            pass
    bi.compute_mahalanobis_distance(alpha=alpha)
    bi.compute_voxels_occupied_in_single_epoch()
    bi.prepare_data_for_export()
    md = bi.merged_df
    store["bi_merged_voxels"] = np.array([list(t) for t in md.index.tolist()], np.int64).reshape(-1, 3)
    for c in md.columns:
        store[f"bi_merged_{c}"] = md[c].to_numpy()
    ex = bi.df
    store["bi_export_voxels"] = np.array([list(t) for t in ex.index.tolist()], np.int64).reshape(-1, 3)
    for c in ex.columns:
        store[f"bi_export_{c}"] = ex[c].to_numpy()


def save(name, store):
    import scipy

    store["meta_versions"] = np.array([f"numpy {np.__version__}", f"pandas {pd.__version__}", f"scipy {scipy.__version__}"])
    os.makedirs(OUT, exist_ok=True)
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **store)
    print(f"{name}: {len(store)} arrays")


def kat_voxelize():
    """KAT-1 boundary values: x on voxel faces, negatives, UTM scale."""
    store = {}
    xs = np.array([0.3, 0.6, 0.7, 0.30000000000000004, 0.1, 0.2, 1.0, 0.0, -0.0, 476850.05, -0.1, -0.3,
                   -0.30000000000000004, 1e-17, -1e-17, 5429100.35, 0.9999999999999999, 123456.7, -98765.4321])
    df = pd.DataFrame({"X": xs, "Y": xs[::-1].copy(), "Z": xs * 0.5})
    for tag, vs, origin in (("a", 0.1, [0, 0, 0]), ("b", 0.25, [0.05, 0.05, 0.05]), ("c", 0.07, [-3.3, 1e3, 0.011])):
        v = new_vapc(df, vs, origin)
        v.voxelize()
        store[f"{tag}_vs"] = np.float64(vs)
        store[f"{tag}_origin"] = np.asarray(origin, np.float64)
        for d in "xyz":
            store[f"{tag}_voxel_{d}"] = v.df[f"voxel_{d}"].to_numpy()
    for c in df.columns:
        store[f"in_{c}"] = df[c].to_numpy()
    save("kat_voxelize", store)


def cloud10():
    pts = np.array([[0.10, 0.10, 0.10, 2], [0.40, 0.20, 0.30, 2], [0.25, 0.45, 0.05, 5], [0.30, 0.30, 0.30, 3],
                    [0.60, 0.10, 0.10, 1], [0.90, 0.40, 0.20, 1], [-0.10, 0.10, 0.10, 7], [0.5, 0.5, 0.5, 4],
                    [0.75, 0.75, 0.75, 4], [0.99, 0.51, 0.6, 6]])
    return pd.DataFrame(pts, columns=["X", "Y", "Z", "classification"])


def kat_cloud10():
    """KAT-2..6: the explicit 10-point cloud, voxel 0.5."""
    df = cloud10()
    store = {}
    mask = pd.DataFrame({"X": [0.7, 1.2], "Y": [0.2, 1.2], "Z": [0.2, 1.2]})
    run_cloud(store, df, 0.5, stats={"classification": ["mode", "mode_count,0.3", "mode_count,0.5", "mean", "min", "max", "sum", "median", "std", "var"]},
              mask_df=mask, buffer_size=1)
    e2 = df[["X", "Y", "Z"]].iloc[:6].to_numpy() + np.array([0.03, -0.02, 0.01])
    e2 = np.vstack([e2, [[0.55, 0.6, 0.7], [0.8, 0.9, 0.6], [0.7, 0.7, 0.9], [1.2, 0.1, 0.1], [1.3, 0.2, 0.2]]])
    run_bitemporal(store, df, pd.DataFrame(e2, columns=["X", "Y", "Z"]), 0.5, alpha=0.005)
    save("kat_cloud10", store)


def plane500():
    """The seeded fixture of the reference's tests/test_compute.py:69-127 (np.random.seed(42))."""
    np.random.seed(42)
    x = np.random.uniform(0.0, 10.0, 500)
    y = np.random.uniform(0.0, 10.0, 500)
    z = np.zeros(500) + np.random.normal(0.0, 0.1, 500)
    inten = np.random.normal(-3, 5, 500)
    nor = np.random.randint(1, 5, 500)
    df = pd.DataFrame({"X": x, "Y": y, "Z": z, "intensity": inten, "number_of_returns": nor.astype(np.float64)})
    red = [120, 200, 3]
    df["X"] -= df["X"].min() - red[0]
    df["Y"] -= df["Y"].min() - red[1]
    df["Z"] -= df["Z"].min() - red[2]
    for vs in (1.0, 0.5, 2.5):
        store = {}
        run_cloud(store, df, vs, stats={"intensity": ["mean", "min", "max", "median", "sum"],
                                        "number_of_returns": ["mode", "mode_count,0.1", "mode_count,0.3"]})
        v = new_vapc(df[["X", "Y", "Z"]], vs)
        v.compute = ["point_count"]
        v.compute_requested_attributes()
        v.filter_attributes("point_count", ">", 3)
        store["filter_point_count_gt3_rows"] = np.int64(v.df.shape[0])
        v = new_vapc(df[["X", "Y", "Z"]], vs)
        v.compute_reduction_point()
        store["reduction_point"] = np.asarray(v.reduction_point, np.int64)
        save(f"plane500_vs{vs}", store)


def uniform_cloud(n, seed, lo, hi, quant=1e-4, nclass=6):
    rng = np.random.default_rng(seed)
    q = [rng.integers(int(l / quant), int(h / quant), n) for l, h in zip(lo, hi)]
    df = pd.DataFrame({"X": q[0] * quant, "Y": q[1] * quant, "Z": q[2] * quant})
    df["classification"] = rng.integers(1, nclass + 1, n).astype(np.float64)
    return df


def uniform5000():
    """5 000 LAS-like points (1e-4 m quantisation), voxel 0.5 and 0.25, with a mask cloud and a
    second epoch (jittered + shifted subset + new points)."""
    df = uniform_cloud(5000, 7, (-3.0, 2.0, 0.0), (7.0, 12.0, 2.0))
    rng = np.random.default_rng(8)
    mask = uniform_cloud(60, 9, (0.0, 4.0, 0.0), (3.0, 8.0, 2.0))
    e2 = df[["X", "Y", "Z"]].to_numpy()[rng.random(5000) < 0.9]
    e2 = e2 + rng.normal(0, 0.01, e2.shape)
    moved = e2[:, 0] > 4.0
    e2[moved] += np.array([0.15, 0.0, 0.05])
    extra = uniform_cloud(300, 10, (7.5, 2.0, 0.0), (9.0, 12.0, 2.0))[["X", "Y", "Z"]].to_numpy()
    e2 = np.vstack([e2, extra])
    for vs in (0.5, 0.25):
        store = {}
        run_cloud(store, df, vs, stats={"classification": ["mode", "mode_count,0.1", "mode_count,0.34", "median", "mean", "var"]},
                  mask_df=mask, buffer_size=1 if vs == 0.5 else 2)
        run_bitemporal(store, df, pd.DataFrame(e2, columns=["X", "Y", "Z"]), vs if vs == 0.5 else 1.0, alpha=0.005 if vs == 0.5 else 0.999)
        save(f"uniform5000_vs{vs}", store)


def shifted_origin():
    """Negative coordinates, non-zero origin, voxel size that is inexact in binary."""
    rng = np.random.default_rng(11)
    n = 3000
    df = pd.DataFrame({"X": rng.integers(-40000, 40000, n) * 1e-3, "Y": rng.integers(-10000, 5000, n) * 1e-3,
                       "Z": rng.integers(-3000, 3000, n) * 1e-3})
    df["classification"] = rng.integers(0, 4, n).astype(np.float64)
    store = {}
    run_cloud(store, df, 0.3, origin=[1.5, -2.25, 0.3], stats={"classification": ["mode", "mode_count,0.25"]})
    save("shifted_origin_vs0.3", store)


def utm_offset():
    """UTM-scale coordinates: the reference's cov_* columns are ill-conditioned here
    (SURVEY 6.2); the fixture documents what it outputs anyway."""
    df = uniform_cloud(4000, 12, (0.0, 0.0, 0.0), (6.0, 6.0, 3.0))
    df["X"] += 476850.0
    df["Y"] += 5429100.0
    df["Z"] += 312.0
    store = {}
    run_cloud(store, df, 0.25)
    save("utm_offset_vs0.25", store)


def grid8():
    """The 8-voxel x 5-point fixture of the reference's tests/test_voxel.py:9-68 (seeded here)."""
    rng = np.random.default_rng(13)
    vx, vy, vz = [0, 1, 1, 3, 4, 5, 6, 9], [0, 1, 2, 3, 5, 2, 5, 7], [0, 1, 0, 0, 1, 0, 2, 1]
    mx, my, mz = [0, 0, 1, 1], [0, 1, 1, 2], [0, 0, 1, 0]

    def mk(ax, ay, az):
        x, y, z = [], [], []
        for a, b, c in zip(ax, ay, az):
            x += list(rng.uniform(a, a + 1, 5))
            y += list(rng.uniform(b, b + 1, 5))
            z += list(rng.uniform(c, c + 1, 5))
        return pd.DataFrame({"X": x, "Y": y, "Z": z})

    store = {}
    run_cloud(store, mk(vx, vy, vz), 1, mask_df=mk(mx, my, mz), buffer_size=1)
    save("grid8_vs1", store)


def real_file():
    """BASELINE configs[0]: the reference's own sample file tests/test_data/vapc_in.laz (LAS 1.2, point format 2,
    LASzip-compressed), voxel 0.25 m.  laspy is not installable offline, so the points are decoded with this
    repository's LAZ reader (vapc_b200/las.py -> vapc_laz_decode), which tests/test_laz.py pins against the file
    header's bounding box and the reference's own known answer for this file ("13.43 percent of the voxel space is
    occupied" at voxel 0.5, tests/integration_tests/test_main.py:216-218).  Every 50th point (10 954 points) keeps the
    fixture small and the reference's per-voxel Python eigen step short; tests/test_data/small_mask.laz is the mask."""
    from vapc_b200 import las

    root = "/root/reference/tests/test_data/"
    _, c = las.read_las(root + "vapc_in.laz")
    sel = slice(0, None, 50)
    df = pd.DataFrame({"X": c["x"][sel], "Y": c["y"][sel], "Z": c["z"][sel], "red": c["red"][sel].astype(np.float64),
                       "classification": (c["red"][sel] // 8000).astype(np.float64)})
    _, m = las.read_las(root + "small_mask.laz")
    mask = pd.DataFrame({"X": m["x"], "Y": m["y"], "Z": m["z"]})
    store = {}
    run_cloud(store, df, 0.25, stats={"red": ["mean", "min", "max", "median"], "classification": ["mode", "mode_count,0.1"]}, mask_df=mask,
              buffer_size=1)
    save("vapc_in_every50th_vs0.25", store)


def clusters_nn():
    """compute_clusters (vapc.py:1274-1357) and merge_vapcs_with_closest_point
    (bi_vapc.py:51-74) of the unmodified reference.
      cl_sparse_*   sparse cloud, default cluster_by (voxel triple), distances 1.0 / 1.8 / 2.0 voxels: every KD-tree query
                    finds fewer than 50 rows within the distance, i.e. the regime where the sweep is not truncated;
      cl_xyz_*      cluster_by = X, Y, Z (float rows), distance 0.3 m;
      cl_blobs_*    the reference's own unit-test fixture (tests/test_compute.py:128-160, 235-242): 3 blobs of 100
                    points, voxel 0.2, distance 10 voxels — the truncated regime (every query returns 50 rows);
      nn_*          two epochs of 4000 / 4300 points, voxel 0.5: prepare_data_for_mahalanobis_distance,
                    merge_vapcs_with_closest_point, compute_mahalanobis_distance."""
    from sklearn.datasets import make_blobs

    store = {}
    rng = np.random.default_rng(77)
    pts = np.round(rng.random((900, 3)) * [12.0, 12.0, 6.0], 4)
    df = pd.DataFrame({"X": pts[:, 0], "Y": pts[:, 1], "Z": pts[:, 2]})
    for c in "XYZ":
        store[f"in_sparse_{c}"] = df[c].to_numpy()
    store["in_sparse_voxel_size"] = np.float64(0.5)
    for d in (1.0, 1.8, 2.0):
        v = new_vapc(df, 0.5)
        v.compute_clusters(cluster_distance=d)
        store[f"cl_sparse_d{d}_cluster_id"] = v.df["cluster_id"].to_numpy()
        store[f"cl_sparse_d{d}_cluster_size"] = v.df["cluster_size"].to_numpy()
        store[f"cl_sparse_d{d}__columns"] = np.array(v.df.columns.tolist())
    v = new_vapc(df, 0.5)
    v.compute_clusters(cluster_distance=0.3, cluster_by=["X", "Y", "Z"])
    store["cl_xyz_d0.3_cluster_id"] = v.df["cluster_id"].to_numpy()
    store["cl_xyz_d0.3_cluster_size"] = v.df["cluster_size"].to_numpy()
    store["cl_xyz_d0.3__columns"] = np.array(v.df.columns.tolist())
    bp, bid = make_blobs(n_samples=300, n_features=3, centers=3, cluster_std=1.0, random_state=42)
    bdf = pd.DataFrame({"X": bp[:, 0], "Y": bp[:, 1], "Z": bp[:, 2], "cluster_id": bid})
    for c in "XYZ":
        store[f"in_blobs_{c}"] = bdf[c].to_numpy()
    store["in_blobs_true_id"] = bid
    v = new_vapc(bdf, 0.2)
    v.compute_clusters(cluster_distance=10.0)
    store["cl_blobs_cluster_id"] = v.df["cluster_id"].to_numpy()
    store["cl_blobs_cluster_size"] = v.df["cluster_size"].to_numpy()
    # two epochs
    e1 = np.round(rng.random((4000, 3)) * [10.0, 10.0, 3.0], 4)
    e2 = np.concatenate([np.round(e1[:3500] + rng.normal(0, 0.03, (3500, 3)), 4), np.round(rng.random((800, 3)) * [10.0, 10.0, 3.0] + [4.0, 0, 0], 4)])
    for k, e in (("1", e1), ("2", e2)):
        for j, c in enumerate("XYZ"):
            store[f"in_nn_e{k}_{c}"] = e[:, j]
    store["in_nn_voxel_size"] = np.float64(0.5)
    a = new_vapc(pd.DataFrame(e1, columns=list("XYZ")), 0.5)
    b = new_vapc(pd.DataFrame(e2, columns=list("XYZ")), 0.5)
    bi = vapc.BiTemporalVapc([a, b])
    bi.prepare_data_for_mahalanobis_distance()
    bi.merge_vapcs_with_closest_point()
    bi.compute_mahalanobis_distance(alpha=0.005)
    md = bi.merged_df
    store["nn__columns"] = np.array(md.columns.tolist())
    for c in md.columns:
        store[f"nn_{c}"] = md[c].to_numpy()
    save("clusters_nn", store)


def label_change():
    """label_by_mask (vapc.py:614-677) and the in-memory core of extract_areas_with_change_using_mahalanobis_distance
    (vapc_tools.py:625-660: two-epoch change -> voxels with change_type >= 1 as the mask -> select_by_mask + label_by_mask of both
    epochs, vapc_tools.py:589-622) of the unmodified reference, frames handed over directly instead of through LAS files.
      lab_unique_*   scene of 3000 points labelled by a mask of unique voxels carrying two attributes;
      lab_dup_*      the same with a mask holding several rows per voxel (the left join multiplies rows);
      chg_*          epochs of 6000 / 6200 points, voxel 0.5, alpha 0.005: the mask frame and, per epoch, the selected + labelled frame."""
    store = {}
    rng = np.random.default_rng(2024)
    scene = uniform_cloud(3000, 31, (0.0, 0.0, 0.0), (8.0, 6.0, 3.0))[["X", "Y", "Z"]].assign(_row=np.arange(3000, dtype=np.float64))
    mpts = uniform_cloud(400, 32, (1.0, 1.0, 0.0), (6.0, 5.0, 2.0))[["X", "Y", "Z"]]
    mpts["label"] = rng.integers(0, 5, len(mpts)).astype(np.float64)
    mpts["score"] = np.round(rng.random(len(mpts)), 3)
    vs = 0.5
    for c in scene.columns:
        store[f"in_scene_{c}"] = scene[c].to_numpy()
    for c in mpts.columns:
        store[f"in_mask_{c}"] = mpts[c].to_numpy()
    store["in_voxel_size"] = np.float64(vs)
    # unique mask voxels: the mask reduced to voxel centres first
    m = new_vapc(mpts, vs, return_at="center_of_voxel")
    m.reduce_to_voxels()
    store["lab_unique_mask__columns"] = np.array(m.df.columns.tolist())
    cols_to(store, "lab_unique_mask_", m.df)
    m = new_vapc(m.df, vs)  # a fresh object over the reduced frame, as point_cloud_to_vapc gives after reading the mask file
    p = new_vapc(scene, vs)
    p.label_by_mask(m, label_attributes=["label", "score"])
    store["lab_unique__columns"] = np.array(p.df.columns.tolist())
    cols_to(store, "lab_unique_", p.df.reset_index(drop=True))
    store["lab_unique_index_names"] = np.array([str(n) for n in p.df.index.names])
    # several mask rows per voxel
    m = new_vapc(mpts, vs)
    p = new_vapc(scene, vs)
    p.label_by_mask(m, label_attributes=["label"])
    store["lab_dup__columns"] = np.array(p.df.columns.tolist())
    cols_to(store, "lab_dup_", p.df.reset_index(drop=True))
    # default arguments: the join refuses the overlapping voxel columns
    try:
        new_vapc(scene, vs).label_by_mask(new_vapc(mpts, vs))
        store["lab_default_raises"] = np.array("")
    except Exception as exc:  # noqa: BLE001
        store["lab_default_raises"] = np.array(type(exc).__name__)
    # change areas
    e1 = np.round(rng.random((6000, 3)) * [12.0, 10.0, 3.0], 4)
    moved = e1[:5200].copy()
    moved[(moved[:, 0] > 5.0) & (moved[:, 0] < 7.0)] += [0.15, 0.0, 0.1]
    e2 = np.concatenate([np.round(moved + rng.normal(0, 0.01, moved.shape), 4), np.round(rng.random((1000, 3)) * [12.0, 10.0, 3.0] + [3.0, 0, 0], 4)])
    alpha = 0.005
    for k, e in (("1", e1), ("2", e2)):
        for j, c in enumerate("XYZ"):
            store[f"in_chg_e{k}_{c}"] = e[:, j]
    store["in_chg_alpha"] = np.float64(alpha)
    frames = [pd.DataFrame(e, columns=list("XYZ")) for e in (e1, e2)]
    bi = vapc.BiTemporalVapc([new_vapc(frames[0], vs), new_vapc(frames[1], vs)])
    bi.prepare_data_for_mahalanobis_distance()
    bi.merge_vapcs_with_same_voxel_index()
    bi.compute_mahalanobis_distance(alpha=alpha)
    bi.compute_voxels_occupied_in_single_epoch()
    bi.prepare_data_for_export()
    bi.df = bi.df[(bi.df["change_type"] >= 1)]
    mask_frame = bi.df.reset_index(drop=True)
    store["chg_mask__columns"] = np.array(mask_frame.columns.tolist())
    cols_to(store, "chg_mask_", mask_frame)
    for k, f in (("1", frames[0]), ("2", frames[1])):
        vm = vapc.Vapc(vs)
        vm.df = mask_frame.copy()
        sc = vapc.Vapc(vs)
        sc.df = f.assign(_row=np.arange(len(f), dtype=np.float64))
        sc.select_by_mask(vm, segment_in_or_out="in")
        sc.label_by_mask(vm, ["mahalanobi_significance", "p_value", "change_type"])
        out = sc.df.reset_index(drop=True)
        store[f"chg_out{k}__columns"] = np.array(out.columns.tolist())
        cols_to(store, f"chg_out{k}_", out)
    save("label_change", store)


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "clusters_nn":
        clusters_nn()
        sys.exit(0)
    if len(sys.argv) > 1 and sys.argv[1] == "label_change":
        label_change()
        sys.exit(0)
    if len(sys.argv) > 1 and sys.argv[1] == "real_file":
        real_file()
        sys.exit(0)
    kat_voxelize()
    kat_cloud10()
    plane500()
    uniform5000()
    shifted_origin()
    utm_offset()
    grid8()
    real_file()
    clusters_nn()
    label_change()
