"""GPU tests of the reference-shaped facade (vapc_b200.Vapc / BiTemporalVapc) against the golden
fixtures the UNMODIFIED reference produced (tests/golden, oracle/make_golden.py), column by column,
and a replay of the reference's own unit-test strategy (known 8-voxel grid, seeded plane, dispatch,
filter known answer).  Tolerances: tests/parity.py."""
import numpy as np
import pandas as pd
import pytest

from tests import parity as P

pytestmark = pytest.mark.gpu

FULL = ["point_count", "point_density", "center_of_gravity", "std_of_cog", "covariance_matrix", "eigenvalues",
        "geometric_features", "distance_to_center_of_gravity", "closest_to_center_of_gravity", "center_of_voxel", "corner_of_voxel"]
MODES = ["center_of_voxel", "corner_of_voxel", "center_of_gravity", "closest_to_center_of_gravity"]


@pytest.fixture(scope="module")
def vapc():
    import vapc_b200

    return vapc_b200


def input_frame(gold):
    cols = [k[3:] for k in gold if k.startswith("in_") and gold[k].ndim == 1 and len(gold[k]) == len(gold["in_X"])
            and not k.startswith(("in_mask_", "in_e1_", "in_e2_"))]
    cols = ["X", "Y", "Z"] + [c for c in cols if c not in "XYZ"]
    return pd.DataFrame({c: gold["in_" + c] for c in cols})


def new(vapc, gold, df=None, **kw):
    origin = gold["in_origin"].tolist()
    v = vapc.Vapc(float(gold["in_voxel_size"]), origin=origin, **kw)
    v.df = (input_frame(gold) if df is None else df).copy()
    v.original_attributes = v.df.columns.tolist()
    return v


@pytest.mark.parametrize("name", P.cloud_cases())
def test_covariance_formula_switch_reproduces_reference_columns(vapc, name):
    """covariance_formula = "reference": the nine cov_** columns of the frame EQUAL the reference's, bit for bit (its raw-moment
    formula over pandas' Kahan group sums) — for users who diff against the reference; the default stays the exact covariance."""
    gold = P.load_golden(name)
    v = new(vapc, gold)
    assert v.covariance_formula == "exact"
    v.covariance_formula = "reference"
    v.compute = ["covariance_matrix", "eigenvalues"]
    v.compute_requested_attributes()
    df = v.df
    for c in ("cov_xx", "cov_xy", "cov_xz", "cov_yx", "cov_yy", "cov_yz", "cov_zx", "cov_zy", "cov_zz"):
        np.testing.assert_array_equal(df[c].to_numpy(), np.real(np.asarray(gold["pt_" + c])), err_msg=f"{name} {c}")
    # the eigenvalues do not follow the switch (the reference takes them from np.cov, vapc.py:1055)
    w = new(vapc, gold)
    w.compute = ["eigenvalues"]
    w.compute_requested_attributes()
    for i in (1, 2, 3):
        np.testing.assert_array_equal(df[f"Eigenvalue_{i}"].to_numpy(), w.df[f"Eigenvalue_{i}"].to_numpy())
    bad = new(vapc, gold)
    bad.covariance_formula = "raw"
    bad.compute = ["covariance_matrix"]
    with pytest.raises(ValueError):
        bad.compute_requested_attributes()


@pytest.mark.parametrize("name", P.cloud_cases())
def test_full_attribute_set_matches_reference(vapc, name):
    gold = P.load_golden(name)
    v = new(vapc, gold)
    v.compute = list(FULL)
    v.compute_requested_attributes()
    df = v.df
    assert df.columns.tolist() == gold["pt__columns"].tolist(), "column names and order"
    assert len(df) == len(gold["in_X"])
    coords = np.stack([gold["in_X"], gold["in_Y"], gold["in_Z"]], 1)
    cmax = np.abs(coords).max()
    vs = float(gold["in_voxel_size"])
    # integers: bit-exact, same dtype
    for c in ("voxel_x", "voxel_y", "voxel_z", "point_count"):
        assert df[c].dtype == np.int64
        np.testing.assert_array_equal(df[c].to_numpy(), gold["pt_" + c], err_msg=c)
    for c in ("center_x", "center_y", "center_z", "corner_x", "corner_y", "corner_z"):
        np.testing.assert_array_equal(df[c].to_numpy(), gold["pt_" + c], err_msg=c)
    np.testing.assert_array_equal(df["point_density"].to_numpy(), gold["pt_point_density"])  # count / voxel_size ** 3, Python's pow
    for d in "xyz":
        P.assert_close_rel(df[f"cog_{d}"].to_numpy(), gold[f"pt_cog_{d}"], what=f"cog_{d}")
        P.assert_close_rel(df[f"std_{d}"].to_numpy(), gold[f"pt_std_{d}"], what=f"std_{d}", atol=8 * P.EPS * cmax)
    # covariance, H1 rule, per point (all points of a voxel carry the voxel's value)
    six = ("cov_xx", "cov_xy", "cov_xz", "cov_yy", "cov_yz", "cov_zz")
    got6 = np.stack([df[c].to_numpy() for c in six], 1)
    ref6 = np.stack([gold["pt_" + c] for c in six], 1)
    cnt = gold["pt_point_count"]
    grp = pd.DataFrame({"k0": gold["pt_voxel_x"], "k1": gold["pt_voxel_y"], "k2": gold["pt_voxel_z"],
                        "s": (coords ** 2).sum(1)}).groupby(["k0", "k1", "k2"])["s"].transform("mean").to_numpy()
    P.assert_cov_h1(got6, ref6, cnt, grp)
    for a, b in (("cov_xy", "cov_yx"), ("cov_xz", "cov_zx"), ("cov_yz", "cov_zy")):
        np.testing.assert_array_equal(df[a].to_numpy(), df[b].to_numpy())
    ref_eig = np.stack([gold[f"pt_Eigenvalue_{i}"] for i in (1, 2, 3)], 1)
    got_eig = np.stack([df[f"Eigenvalue_{i}"].to_numpy() for i in (1, 2, 3)], 1)
    floor = P.noise_floor(cmax)
    P.assert_eig_h7(got_eig, ref_eig, cnt, abs_floor=floor)
    P.assert_features_h7(np.stack([df[c].to_numpy() for c in P.FEATURES], 1), np.stack([gold["pt_" + c] for c in P.FEATURES], 1), ref_eig, cnt,
                         abs_floor=floor)
    with np.errstate(invalid="ignore"):
        ok = (cnt > 1) & (ref_eig[:, 0] > 1e9 * floor)
    for i in (1, 2, 3):
        assert np.all(np.abs(df[f"Eigenvalue_{i}_n"].to_numpy()[ok] - gold[f"pt_Eigenvalue_{i}_n"][ok]) <= 1e-9)
    tol = 8 * P.EPS * max(vs, cmax)
    assert np.all(np.abs(df["distance"].to_numpy() - gold["pt_distance"]) <= 1e-9 * gold["pt_distance"] + tol)
    assert np.all(np.abs(df["min_distance"].to_numpy() - gold["pt_min_distance"]) <= 1e-9 * gold["pt_min_distance"] + tol)
    v2 = new(vapc, gold)
    v2.compute_percentage_occupied()
    assert v2.percentage_occupied == float(gold["percentage_occupied"])


def check_closest_h5(out, gold):
    """Closest-to-centroid reduction against the reference's frame by the H5 rule: the reference keeps every
    point with distance == min_distance; its distances carry the rounding of ITS centroid, so per voxel
      * every row we return must be one of the reference's near-tied points (distance <= min * (1 + 1e-12) + ulp scale),
      * where the runner-up is clearly farther (the overwhelming majority) our rows must be the reference's rows exactly,
        every column included (the other columns carry the winner's own attribute values).
    Returns the number of near-tie voxels (printed by the caller's -s run)."""
    vs, origin = float(gold["in_voxel_size"]), gold["in_origin"]
    pts = np.stack([gold["in_X"], gold["in_Y"], gold["in_Z"]], 1)
    vox = np.stack([gold["pt_voxel_x"], gold["pt_voxel_y"], gold["pt_voxel_z"]], 1)
    dist, dmin = gold["pt_distance"], gold["pt_min_distance"]
    uniq, inv = np.unique(vox, axis=0, return_inverse=True)
    inv = inv.ravel()
    scale = 8 * P.EPS * max(vs, np.abs(pts).max())
    near = dist <= dmin * (1 + 1e-12) + scale
    # runner-up distance per voxel among the points that are not near-tied
    far = np.where(near, np.inf, dist)
    second = np.full(len(uniq), np.inf)
    np.minimum.at(second, inv, far)
    vmin = np.zeros(len(uniq))
    vmin[inv] = dmin
    near_count = np.bincount(inv, weights=near.astype(np.float64), minlength=len(uniq))
    clear_voxel = (near_count == 1) & (second > vmin * (1 + 1e-9) + 1e-12)  # one candidate, runner-up clearly farther

    def voxel_row(xyz):
        v = np.floor((xyz - origin) / vs).astype(np.int64)
        lookup = {tuple(r): i for i, r in enumerate(uniq.tolist())}
        return np.array([lookup[tuple(r)] for r in v.tolist()], dtype=np.int64)

    cols = out.columns.tolist()
    assert cols == gold["red_closest_to_center_of_gravity__columns"].tolist()
    got = out.to_numpy()
    ref = np.stack([gold[f"red_closest_to_center_of_gravity_{c}"] for c in cols], 1)
    gv, rv = voxel_row(got[:, :3]), voxel_row(ref[:, :3])
    assert set(gv.tolist()) == set(range(len(uniq))), "one row (at least) for every voxel"
    # every returned point is a near-tied point of its voxel in the reference
    pt_lookup = {}
    for i in np.nonzero(near)[0]:
        pt_lookup.setdefault((inv[i],) + tuple(pts[i].tolist()), i)
    for r, vrow in zip(got[:, :3].tolist(), gv.tolist()):
        assert (vrow,) + tuple(r) in pt_lookup, "returned point is not in the reference's tied set of its voxel"
    # clear winners: identical rows, all columns, same (X, Y, Z) order
    a, b = got[clear_voxel[gv]], ref[clear_voxel[rv]]
    np.testing.assert_array_equal(a, b)
    return int((~clear_voxel).sum())


@pytest.mark.parametrize("name", P.cloud_cases())
@pytest.mark.parametrize("mode", MODES)
def test_reduce_to_voxels_matches_reference(vapc, name, mode):
    gold = P.load_golden(name)
    v = new(vapc, gold, return_at=mode)
    v.reduce_to_voxels()
    assert v.reduced is True
    assert v.df.columns.tolist() == gold[f"red_{mode}__columns"].tolist()
    ref = np.stack([gold[f"red_{mode}_{c}"] for c in "XYZ"], 1)
    got = v.df[["X", "Y", "Z"]].to_numpy()
    if mode in ("center_of_voxel", "corner_of_voxel"):
        np.testing.assert_array_equal(got, ref)  # exact values, exact (lexicographic voxel) order
    elif mode == "center_of_gravity":
        assert got.shape == ref.shape
        # rows are sorted by the centroid floats; a last-bit difference in X may swap neighbours whose X agree
        # to the last bit, so align the rows by the voxel each centroid lies in
        def by_voxel(a):
            vox = np.floor((a - gold["in_origin"]) / float(gold["in_voxel_size"])).astype(np.int64)
            return a[np.lexsort((vox[:, 2], vox[:, 1], vox[:, 0]))]
        np.testing.assert_allclose(by_voxel(got), by_voxel(ref), rtol=1e-9, atol=0)
    else:
        check_closest_h5(v.df, gold)
    others = [c for c in v.df.columns if c not in "XYZ"]
    if mode != "closest_to_center_of_gravity" and others:  # (closest: all columns are compared inside check_closest_h5)
        if mode == "center_of_gravity":
            def vox_order(a):
                vox = np.floor((a - gold["in_origin"]) / float(gold["in_voxel_size"])).astype(np.int64)
                return np.lexsort((vox[:, 2], vox[:, 1], vox[:, 0]))
            a = v.df[others].to_numpy()[vox_order(got)]
            b = np.stack([gold[f"red_{mode}_{c}"] for c in others], 1)[vox_order(ref)]
            np.testing.assert_array_equal(a, b)
        else:
            np.testing.assert_array_equal(v.df[others].to_numpy(), np.stack([gold[f"red_{mode}_{c}"] for c in others], 1))


@pytest.mark.parametrize("name", P.cloud_cases())
def test_centroid_columns_and_reduced_frames_equal_reference_bit_for_bit(vapc, name):
    """centroid_formula = "reference" (the facade's default): the centroid is pandas' groupby mean to the bit, so cog_*, distance and
    min_distance EQUAL the golden columns, and the centre-of-gravity / closest-to-centroid reductions return the reference's frames
    row for row — same rows (every exactly tied point included), same order, every column: nothing left to the H5 tolerance.  With
    "exact" (the reduction's own mean, no extra pass) the H5 rule holds as before."""
    gold = P.load_golden(name)
    v = new(vapc, gold)
    assert v.centroid_formula == "reference"
    v.compute = ["center_of_gravity", "distance_to_center_of_gravity", "closest_to_center_of_gravity"]
    v.compute_requested_attributes()
    df = v.df
    for c in ("cog_x", "cog_y", "cog_z", "distance", "min_distance"):
        np.testing.assert_array_equal(df[c].to_numpy(), gold["pt_" + c], err_msg=f"{name} {c}")
    for mode in ("center_of_gravity", "closest_to_center_of_gravity"):
        w = new(vapc, gold, return_at=mode)
        w.reduce_to_voxels()
        cols = gold[f"red_{mode}__columns"].tolist()
        assert w.df.columns.tolist() == cols
        ref = np.stack([gold[f"red_{mode}_{c}"] for c in cols], 1)
        np.testing.assert_array_equal(w.df.to_numpy(), ref, err_msg=f"{name} {mode}")
    s_ref = new(vapc, gold)
    assert s_ref.std_formula == "exact"
    s_ref.std_formula = "reference"
    s_ref.compute_std_of_cog()
    for c in ("std_x", "std_y", "std_z"):
        np.testing.assert_array_equal(s_ref.df[c].to_numpy(), gold["pt_" + c], err_msg=f"{name} {c}")
    s_bad = new(vapc, gold)
    s_bad.std_formula = "welford"
    with pytest.raises(ValueError):
        s_bad.compute_std_of_cog()
    e = new(vapc, gold, return_at="closest_to_center_of_gravity")
    e.centroid_formula = "exact"
    e.reduce_to_voxels()
    check_closest_h5(e.df, gold)
    bad = new(vapc, gold)
    bad.centroid_formula = "kahan"
    with pytest.raises(ValueError):
        bad.compute_center_of_gravity()


@pytest.mark.parametrize("name", [n for n in P.cloud_cases() if any(k.startswith("stat_") for k in P.load_golden(n))])
def test_attribute_statistics_match_reference(vapc, name):
    gold = P.load_golden(name)
    keys = [k for k in gold if k.startswith("stat_") and not k.startswith("stat_voxel_")]
    attrs = sorted((k[3:] for k in gold if k.startswith("in_")), key=len, reverse=True)
    spec = {}
    for k in keys:
        col = k[5:]
        attr = next(a for a in attrs if col.startswith(a + "_"))
        spec.setdefault(attr, []).append(col[len(attr) + 1:])
    v = new(vapc, gold, attributes=spec)
    v.compute_requested_statistics_per_attributes()
    assert v.attributes_per_voxel is True
    for k in keys:
        col = k[5:]
        got, ref = v.df[col].to_numpy(), gold[k]
        # every statistic equals the reference's column: the integer ones by construction, mean / sum / var / std because the kernel
        # adds a voxel's values the way numpy adds the reference's slice (tests/test_gpu_kernels.py::test_attribute_statistics_vs_oracle)
        np.testing.assert_array_equal(got, ref, err_msg=col)
        if "mode" in col:
            assert got.dtype == np.int64


@pytest.mark.parametrize("name", [n for n in P.cloud_cases() if "mask_in_rows" in P.load_golden(n)])
def test_mask_and_buffer_match_reference(vapc, name):
    gold = P.load_golden(name)
    mask_df = pd.DataFrame({c: gold["in_mask_" + c] for c in "XYZ"})
    for mode in ("in", "out"):
        p = new(vapc, gold, df=input_frame(gold).assign(_row=np.arange(len(gold["in_X"]), dtype=np.float64)))
        m = new(vapc, gold, df=mask_df)
        p.select_by_mask(m, segment_in_or_out=mode)
        np.testing.assert_array_equal(p.df["_row"].to_numpy().astype(np.int64), gold[f"mask_{mode}_rows"])
        assert p.voxelized is False and "voxel_x" not in p.df.columns
        assert list(p.df.index.names) == ["idx_voxel_x", "idx_voxel_y", "idx_voxel_z"]
    m = new(vapc, gold, df=mask_df)
    buf = m.compute_voxel_buffer(buffer_size=int(gold["in_mask_buffer_size"]))
    np.testing.assert_array_equal(buf[["voxel_x", "voxel_y", "voxel_z"]].to_numpy(), gold["mask_buffer_voxels"])
    np.testing.assert_array_equal(buf[["X", "Y", "Z"]].to_numpy(), gold["mask_buffer_xyz"])
    p = new(vapc, gold, df=input_frame(gold).assign(_row=np.arange(len(gold["in_X"]), dtype=np.float64)))
    m.voxel_index = False
    m.df = m.buffer_df
    p.select_by_mask(m, segment_in_or_out="in")
    np.testing.assert_array_equal(p.df["_row"].to_numpy().astype(np.int64), gold["mask_buffer_in_rows"])
    with pytest.raises(ValueError):
        new(vapc, gold).select_by_mask(new(vapc, gold, df=mask_df), segment_in_or_out="sideways")


@pytest.mark.parametrize("name", [n for n in P.cloud_cases() if "bi_merged_voxels" in P.load_golden(n)])
def test_bitemporal_matches_reference(vapc, name):
    gold = P.load_golden(name)
    vs, alpha = float(gold["in_bi_voxel_size"]), float(gold["in_bi_alpha"])
    eps = []
    for k in "12":
        v = vapc.Vapc(vs)
        v.df = pd.DataFrame({c: gold[f"in_e{k}_{c}"] for c in "XYZ"})
        v.original_attributes = v.df.columns.tolist()
        eps.append(v)
    bi = vapc.BiTemporalVapc(eps)
    with pytest.raises(ValueError):
        bi.compute_voxels_occupied_in_single_epoch()
    bi.prepare_data_for_mahalanobis_distance()
    bi.merge_vapcs_with_same_voxel_index()
    bi.compute_distance()
    bi.compute_mahalanobis_distance(alpha=alpha)
    bi.compute_voxels_occupied_in_single_epoch()
    bi.prepare_data_for_export()
    md = bi.merged_df
    ref_vox = gold["bi_merged_voxels"]
    got_vox = np.array([list(t) for t in md.index.tolist()], np.int64).reshape(-1, 3)
    # the merged rows follow the epoch-1 centroid sort; a last-bit centroid difference may swap equal-X
    # neighbours, so align by voxel triple
    def order(vox):
        return np.lexsort((vox[:, 2], vox[:, 1], vox[:, 0]))
    og, orf = order(got_vox), order(ref_vox)
    np.testing.assert_array_equal(got_vox[og], ref_vox[orf])
    assert (got_vox == ref_vox).mean() > 0.99
    P.assert_close_rel(md["distance"].to_numpy()[og], gold["bi_merged_distance"][orf], rtol=1e-9, atol=1e-13, what="distance")
    # with the reference's centroid double (centroid_formula = "reference", the default) the joined rows come in the reference's order and
    # the centroids and their distance are the reference's numbers
    np.testing.assert_array_equal(got_vox, ref_vox)
    for c in ("cog_x_x", "cog_y_x", "cog_z_x", "cog_x_y", "cog_y_y", "cog_z_y", "distance"):
        np.testing.assert_array_equal(md[c].to_numpy(), gold["bi_merged_" + c], err_msg=c)
    six = ("cov_xx", "cov_xy", "cov_xz", "cov_yy", "cov_yz", "cov_zz")
    nine = ("cov_xx", "cov_xy", "cov_xz", "cov_yx", "cov_yy", "cov_yz", "cov_zx", "cov_zy", "cov_zz")
    eye = np.eye(3) * 1e-10
    d1, d2 = bi.mahalanobis_d
    # conditioning-aware rule: the reference's own raw-moment covariances carry (|x|/sigma)^2 eps (H1)
    cov_y = np.stack([gold[f"bi_merged_{c}_y"] for c in nine], 1).reshape(-1, 3, 3)[orf]
    cov_x = np.stack([gold[f"bi_merged_{c}_x"] for c in nine], 1).reshape(-1, 3, 3)[orf]
    ssq_x = sum(gold[f"bi_merged_cog_{d}_x"][orf] ** 2 for d in "xyz") + np.trace(cov_x, axis1=1, axis2=2)
    ssq_y = sum(gold[f"bi_merged_cog_{d}_y"][orf] ** 2 for d in "xyz") + np.trace(cov_y, axis1=1, axis2=2)
    rel_x, rel_y = P.h1_cov_rel(cov_x, ssq_x), P.h1_cov_rel(cov_y, ssq_y)
    cov_x, cov_y = cov_x + eye, cov_y + eye
    P.assert_mahalanobis(d1[og], gold["bi_d1"][orf], cov_y, cov_rel=rel_y, what="d1")
    P.assert_mahalanobis(d2[og], gold["bi_d2"][orf], cov_x, cov_rel=rel_x, what="d2")
    p, pr = md["p_value"].to_numpy()[og], gold["bi_merged_p_value"][orf]
    assert np.array_equal(np.isnan(p), np.isnan(pr))
    cond = np.full(len(p), np.inf)
    okc = ~np.isnan(gold["bi_d1"][orf]) & ~np.isnan(gold["bi_d2"][orf])
    cond[okc] = np.maximum(np.linalg.cond(cov_x[okc]), np.linalg.cond(cov_y[okc]))
    well = cond < 1e2
    assert np.all(np.abs(p[well] - pr[well]) <= 1e-15 + 1e-9 * np.abs(pr[well])), float(np.max(np.abs(p[well] - pr[well]) / (1e-15 + 1e-9 * np.abs(pr[well]))))
    stable = (well & (np.abs(pr - alpha) > 1e-6)) | (pr == 0)
    sig, ct = md["mahalanobi_significance"].to_numpy()[og], md["change_type"].to_numpy()[og]
    np.testing.assert_array_equal(sig[stable], gold["bi_merged_mahalanobi_significance"][orf][stable])
    np.testing.assert_array_equal(ct[stable], gold["bi_merged_change_type"][orf][stable])
    # appear / disappear rows and the export layout
    ex = bi.df
    assert ex.columns.tolist() == ["X", "Y", "Z", "distance", "mahalanobi_significance", "p_value", "change_type"]
    n_m = len(md)
    ex_vox = np.array([list(t) for t in ex.index.tolist()], np.int64).reshape(-1, 3)
    np.testing.assert_array_equal(ex_vox[n_m:], gold["bi_export_voxels"][n_m:])
    np.testing.assert_array_equal(ex["change_type"].to_numpy()[n_m:], gold["bi_export_change_type"][n_m:])
    np.testing.assert_allclose(ex[["X", "Y", "Z"]].to_numpy()[n_m:], np.stack([gold[f"bi_export_{c}"] for c in "XYZ"], 1)[n_m:], rtol=1e-9)
    assert np.isnan(ex["distance"].to_numpy()[n_m:]).all()


# ---- the reference's own unit-test strategy, re-expressed against the new package ----
def test_known_voxel_grid(vapc):
    """tests/test_voxel.py:88-229 — 8 known voxels x 5 random points, mask of 4 voxels, voxel size 1."""
    rng = np.random.default_rng(123)
    vox = np.array([[0, 0, 0], [1, 1, 1], [1, 2, 0], [3, 3, 0], [4, 5, 1], [5, 2, 0], [6, 5, 2], [9, 7, 1]])
    pts = np.repeat(vox, 5, axis=0) + rng.uniform(0, 1, (40, 3))
    df = pd.DataFrame(pts, columns=["X", "Y", "Z"])

    def mk(d=df):
        v = vapc.Vapc(voxel_size=1)
        v.df = d.copy()
        v.original_attributes = v.df.columns.tolist()
        return v

    v = mk()
    v.voxelize()
    assert v.voxelized is True
    np.testing.assert_array_equal(np.unique(v.df[["voxel_x", "voxel_y", "voxel_z"]].to_numpy(), axis=0), vox)
    v = mk()
    v.compute_voxel_index()
    assert v.voxel_index is True
    np.testing.assert_array_equal(v.df.index.to_frame(index=False).to_numpy(), np.repeat(vox, 5, axis=0))
    v = mk()
    v.compute_corner_of_voxel()
    np.testing.assert_array_equal(v.df[["corner_x", "corner_y", "corner_z"]].to_numpy(), np.repeat(vox, 5, axis=0))
    v = mk()
    v.return_at = "center_of_voxel"
    v.reduce_to_voxels()
    np.testing.assert_array_equal(v.df[["X", "Y", "Z"]].to_numpy(), vox + 0.5)
    v = mk()
    v.return_at = "center_of_gravity"
    v.reduce_to_voxels()
    means = pts.reshape(8, 5, 3).mean(1)
    got = v.df[["X", "Y", "Z"]].to_numpy()
    np.testing.assert_allclose(got[np.lexsort(got.T[::-1])], means[np.lexsort(means.T[::-1])])
    v = mk()
    v.return_at = "closest_to_center_of_gravity"
    v.reduce_to_voxels()
    d = np.linalg.norm(pts.reshape(8, 5, 3) - means[:, None, :], axis=2)
    want = pts.reshape(8, 5, 3)[np.arange(8), d.argmin(1)]
    got = v.df[["X", "Y", "Z"]].to_numpy()
    np.testing.assert_allclose(got[np.lexsort(got.T[::-1])], want[np.lexsort(want.T[::-1])])
    # mask in / out
    mvox = np.array([[0, 0, 0], [0, 1, 0], [1, 1, 1], [1, 2, 0]])
    mdf = pd.DataFrame(np.repeat(mvox, 5, axis=0) + rng.uniform(0, 1, (20, 3)), columns=["X", "Y", "Z"])
    inside = {tuple(r) for r in mvox.tolist()} & {tuple(r) for r in vox.tolist()}
    for mode, expect in (("in", len(inside) * 5), ("out", (8 - len(inside)) * 5)):
        p, m = mk(), mk(mdf)
        p.select_by_mask(m, segment_in_or_out=mode)
        assert len(p.df) == expect
    from itertools import product
    v = mk()
    v.compute_voxel_buffer(buffer_size=1)
    want = np.unique((vox[:, None] + np.array(list(product(range(-1, 2), repeat=3)))).reshape(-1, 3), axis=0)
    np.testing.assert_array_equal(np.unique(v.buffer_df[["voxel_x", "voxel_y", "voxel_z"]].to_numpy(), axis=0), want)


def test_seeded_plane_known_answers(vapc):
    """tests/test_compute.py:202-232, 246-271, 336-351 on the seeded plane fixture (plane500 golden)."""
    gold = P.load_golden("plane500_vs1.0")
    v = new(vapc, gold, df=input_frame(gold)[["X", "Y", "Z"]])
    v.compute_reduction_point()
    assert v.reduction_point == [120, 200, 3]
    assert v.offset_applied is False
    v.compute_offset()
    assert v.offset_applied is True and v.df["X"].min() == 0 and v.df["Y"].min() == 0 and v.df["Z"].min() == 0
    v.compute_offset()
    assert v.offset_applied is False and v.df["X"].min() == 120 and v.df["Z"].min() == 3
    v = new(vapc, gold, df=input_frame(gold)[["X", "Y", "Z"]])
    v.voxelize()
    v.compute = ["point_count"]
    v.compute_requested_attributes()
    v.filter_attributes(filter_attribute="point_count", min_max_eq=">", filter_value=3)
    assert v.df.shape[0] == 442
    # dispatcher: each name calls compute_<name> exactly once
    from unittest.mock import patch
    for name in vapc.Vapc.AVAILABLE_COMPUTATIONS:
        v = new(vapc, gold, df=input_frame(gold)[["X", "Y", "Z"]])
        v.voxelize()
        v.compute = [name]
        with patch(f"vapc_b200.Vapc.compute_{name}") as mock:
            v.compute_requested_attributes()
            mock.assert_called_once()


def test_edit_through_the_callers_own_reference(vapc):
    """The caller assigns a frame, keeps the reference, and edits it in place WITHOUT reading `.df`: a shifted column changes the
    fingerprint of X / Y / Z that `_get_cloud` compares, a single-cell edit is announced with `invalidate()`."""
    gold = P.load_golden("uniform5000_vs0.25")
    vs = float(gold["in_voxel_size"])
    origin = gold["in_origin"].tolist()

    def expected_counts(frame):
        vox = np.floor((frame[["X", "Y", "Z"]].to_numpy() - np.asarray(origin)) / vs).astype(np.int64)
        _, inv, cnt = np.unique(vox, axis=0, return_inverse=True, return_counts=True)
        return cnt[inv.ravel()]

    mine = input_frame(gold).copy()
    v = vapc.Vapc(vs, origin=origin)
    v.df = mine
    v.compute_point_count()
    assert v._lazy and not v._exposed  # computed, never handed out
    mine["X"] = mine["X"].to_numpy() + 3.0 * vs * (np.arange(len(mine)) % 3)
    v.point_count = v.voxelized = False
    v.compute_point_count()
    np.testing.assert_array_equal(v.df["point_count"].to_numpy(), expected_counts(mine))
    # one cell, between two sampled rows of the fingerprint
    w = vapc.Vapc(vs, origin=origin)
    mine2 = input_frame(gold).copy()
    w.df = mine2
    w.compute_point_count()
    mine2.loc[mine2.index[1], "Z"] = mine2["Z"].iloc[1] + 1000.0
    w.invalidate()
    w.point_count = w.voxelized = False
    w.compute_point_count()
    got = w.df["point_count"].to_numpy()
    np.testing.assert_array_equal(got, expected_counts(mine2))
    assert got[1] == 1


# ---- lazily materialised frames, label_by_mask, the change-area flow ------------------------------------------------------------------
def test_lazy_and_materialised_paths_agree(vapc):
    """Computed columns stay on the device until `.df` is read.  The frame built then, and the reduced frame built from the V-row
    table without ever materialising the per-point columns, must equal what the materialised (host column) path gives."""
    gold = P.load_golden("uniform5000_vs0.25")
    for mode in MODES:
        a = new(vapc, gold, return_at=mode)
        a.compute = list(FULL)
        a.compute_requested_attributes()
        assert a._lazy, "columns are pending on the device"
        a.reduce_to_voxels()  # device path: pending columns -> V rows
        b = new(vapc, gold, return_at=mode)
        b.compute = list(FULL)
        b.compute_requested_attributes()
        per_point = b.df  # materialises every pending column, hands the frame out
        assert not b._lazy and per_point.shape[0] == len(gold["in_X"])
        b.reduce_to_voxels()  # host path: the frame's own columns, cloud rebuilt
        assert a.df.columns.tolist() == b.df.columns.tolist()
        np.testing.assert_array_equal(a.df.to_numpy(), b.df.to_numpy(), err_msg=mode)
    # an in-place edit of the frame that was handed out is honoured (the device copy is dropped on every read of .df)
    v = new(vapc, gold)
    v.compute_point_count()
    frame = v.df
    old_count = frame["point_count"].to_numpy().copy()
    frame.loc[:, "X"] = frame["X"].to_numpy() + 100.0 * (np.arange(len(frame)) % 2)
    v.point_count = False
    v.compute_point_count()  # voxelized is still True: like the reference, grouped by the frame's (now stale) voxel columns
    np.testing.assert_array_equal(v.df["point_count"].to_numpy(), old_count)
    v.point_count = v.voxelized = False
    v.compute_point_count()  # re-voxelised from the edited coordinates
    vs = float(gold["in_voxel_size"])
    pts = v.df[["X", "Y", "Z"]].to_numpy()
    vox = np.floor(pts / vs).astype(np.int64)
    np.testing.assert_array_equal(v.df[["voxel_x", "voxel_y", "voxel_z"]].to_numpy(), vox)
    _, inv, cnt = np.unique(vox, axis=0, return_inverse=True, return_counts=True)
    np.testing.assert_array_equal(v.df["point_count"].to_numpy(), cnt[inv.ravel()])
    # voxel columns assigned by the caller are the grouping (two points share a voxel because the caller says so)
    w = new(vapc, gold)
    w.voxelize()
    f = w.df
    f.loc[:, "voxel_x"] = 7
    f.loc[:, "voxel_y"] = np.arange(len(f)) // 2
    f.loc[:, "voxel_z"] = 0
    w.compute_point_count()
    w.compute_center_of_gravity()
    got = w.df
    np.testing.assert_array_equal(got["point_count"].to_numpy(), np.full(len(got), 2))
    x = got["X"].to_numpy()
    np.testing.assert_allclose(got["cog_x"].to_numpy(), np.repeat((x[0::2] + x[1::2]) / 2, 2), rtol=1e-12)


def test_filter_on_pending_column_equals_host_filter(vapc):
    gold = P.load_golden("uniform5000_vs0.5")
    for op, val in ((">", 3), ("<=", 2), ("==", 4), ("!=", 1)):
        a = new(vapc, gold)
        a.compute = ["point_count", "center_of_gravity"]
        a.compute_requested_attributes()
        assert "point_count" in a._lazy
        a.filter_attributes("point_count", op, val)  # compared and compacted on the device
        b = new(vapc, gold)
        b.compute = ["point_count", "center_of_gravity"]
        b.compute_requested_attributes()
        _ = b.df
        b.filter_attributes("point_count", op, val)  # pandas boolean indexing
        assert a.df.columns.tolist() == b.df.columns.tolist()
        assert a.df.index.equals(b.df.index)
        np.testing.assert_array_equal(a.df.to_numpy(), b.df.to_numpy(), err_msg=op)
    with pytest.raises(KeyError):
        new(vapc, gold).filter_attributes("nope", ">", 1)
    with pytest.raises(ValueError):
        new(vapc, gold).filter_attributes("X", "approximately", 1)


def _frame(gold, prefix, cols=None):
    cols = cols if cols is not None else [k[len(prefix):] for k in gold if k.startswith(prefix) and not k.endswith("__columns")]
    return pd.DataFrame({c: gold[prefix + c] for c in cols})


def test_label_by_mask_matches_reference(vapc):
    """vapc.py:614-677 against the unmodified reference (tests/golden/label_change.npz): unique mask voxels (GPU join), repeated
    mask voxels (the left join multiplies rows), and the default arguments (pandas refuses the overlapping voxel columns)."""
    gold = P.load_golden("label_change")
    vs = float(gold["in_voxel_size"])
    scene = _frame(gold, "in_scene_", ["X", "Y", "Z", "_row"])
    mask_pts = _frame(gold, "in_mask_", ["X", "Y", "Z", "label", "score"])

    def mk(df, **kw):
        v = vapc.Vapc(vs, **kw)
        v.df = df.copy()
        v.original_attributes = v.df.columns.tolist()
        return v

    m = mk(mask_pts, return_at="center_of_voxel")
    m.reduce_to_voxels()
    assert m.df.columns.tolist() == gold["lab_unique_mask__columns"].tolist()
    np.testing.assert_array_equal(m.df.to_numpy(), _frame(gold, "lab_unique_mask_", m.df.columns.tolist()).to_numpy())
    m = mk(m.df)
    p = mk(scene)
    p.label_by_mask(m, label_attributes=["label", "score"])
    out = p.df
    assert out.columns.tolist() == gold["lab_unique__columns"].tolist()
    assert [str(n) for n in out.index.names] == gold["lab_unique_index_names"].tolist()
    assert p.voxelized is False
    ref = _frame(gold, "lab_unique_", out.columns.tolist())
    np.testing.assert_array_equal(out.reset_index(drop=True).to_numpy(), ref.to_numpy())  # NaN where the voxel is not in the mask
    for c in out.columns:
        assert out[c].dtype == ref[c].dtype, c
    p = mk(scene)
    p.label_by_mask(mk(mask_pts), label_attributes=["label"])
    out = p.df.reset_index(drop=True)
    assert out.columns.tolist() == gold["lab_dup__columns"].tolist()
    np.testing.assert_array_equal(out.to_numpy(), _frame(gold, "lab_dup_", out.columns.tolist()).to_numpy())
    assert str(gold["lab_default_raises"]) == "ValueError"
    with pytest.raises(ValueError):
        mk(scene).label_by_mask(mk(mask_pts))


def test_change_area_flow_matches_reference(vapc, tmp_path):
    """The in-memory core of extract_areas_with_change_using_mahalanobis_distance (vapc_tools.py:625-660) against the unmodified
    reference: change voxels (change_type >= 1) as the mask, both epochs clipped to it and labelled; then the file-to-file tool."""
    gold = P.load_golden("label_change")
    vs, alpha = float(gold["in_voxel_size"]), float(gold["in_chg_alpha"])
    frames = [pd.DataFrame({c: gold[f"in_chg_e{k}_{c}"] for c in "XYZ"}) for k in "12"]
    eps = []
    for f in frames:
        v = vapc.Vapc(vs)
        v.df = f.copy()
        v.original_attributes = v.df.columns.tolist()
        eps.append(v)
    bi = vapc.BiTemporalVapc(eps)
    bi.prepare_data_for_mahalanobis_distance()
    bi.merge_vapcs_with_same_voxel_index()
    bi.compute_mahalanobis_distance(alpha=alpha)
    bi.compute_voxels_occupied_in_single_epoch()
    bi.prepare_data_for_export()
    bi.df = bi.df[(bi.df["change_type"] >= 1)]
    mask_frame = bi.df.reset_index(drop=True)
    ref_mask = _frame(gold, "chg_mask_", gold["chg_mask__columns"].tolist())
    assert mask_frame.columns.tolist() == ref_mask.columns.tolist()

    def by_voxel(df):
        vox = np.floor(df[["X", "Y", "Z"]].to_numpy() / vs).astype(np.int64)
        return df.iloc[np.lexsort((vox[:, 2], vox[:, 1], vox[:, 0]))].reset_index(drop=True), vox[np.lexsort((vox[:, 2], vox[:, 1], vox[:, 0]))]

    (gm, gvox), (rm, rvox) = by_voxel(mask_frame), by_voxel(ref_mask)
    # classes are exact away from the alpha threshold: compare the voxel sets per change_type, then the columns on the common rows
    stable = np.isnan(rm["p_value"].to_numpy()) | (np.abs(rm["p_value"].to_numpy() - alpha) > 1e-6)
    ref_keys = {tuple(v) + (int(c),) for v, c, s in zip(rvox.tolist(), rm["change_type"].to_numpy(), stable) if s}
    got_keys = {tuple(v) + (int(c),) for v, c in zip(gvox.tolist(), gm["change_type"].to_numpy())}
    assert ref_keys <= got_keys
    assert len(got_keys - ref_keys) <= int((~stable).sum())
    if stable.all():
        np.testing.assert_array_equal(gvox, rvox)
        np.testing.assert_allclose(gm[["X", "Y", "Z"]].to_numpy(), rm[["X", "Y", "Z"]].to_numpy(), rtol=1e-9)
    # clip + label both epochs with the REFERENCE's mask frame (so that the comparison does not inherit last-bit centroid noise)
    for k, f in zip("12", frames):
        vm = vapc.Vapc(vs)
        vm.df = ref_mask.copy()
        sc = vapc.Vapc(vs)
        sc.df = f.assign(_row=np.arange(len(f), dtype=np.float64))
        sc.select_by_mask(vm, segment_in_or_out="in")
        sc.label_by_mask(vm, ["mahalanobi_significance", "p_value", "change_type"])
        out = sc.df.reset_index(drop=True)
        ref = _frame(gold, f"chg_out{k}_", gold[f"chg_out{k}__columns"].tolist())
        assert out.columns.tolist() == ref.columns.tolist()
        np.testing.assert_array_equal(out.to_numpy(), ref.to_numpy())
        for c in out.columns:
            assert out[c].dtype == ref[c].dtype, c
    # file to file: LAS in, mask LAS + two labelled LAS out (coordinates re-quantised by the LAS writer, attributes stored as float32)
    from vapc_b200 import vapc_tools as T
    from vapc_b200.datahandler import DataHandler

    paths = []
    for k, f in zip("12", frames):
        dh = DataHandler("")
        dh.df = f.copy()
        paths.append(str(tmp_path / f"epoch{k}.las"))
        dh.save_as_las(paths[-1])
    outs = [str(tmp_path / "out1.las"), str(tmp_path / "out2.las")]
    res = T.extract_areas_with_change_using_mahalanobis_distance(paths[0], paths[1], str(tmp_path / "mask.las"), outs[0], outs[1], vs, alpha)
    assert res is not False
    for k, o in zip("12", outs):
        dh = DataHandler(o)
        dh.load_las_files()
        ref = _frame(gold, f"chg_out{k}_", gold[f"chg_out{k}__columns"].tolist())
        assert {"mahalanobi_significance", "p_value", "change_type"} <= set(dh.df.columns)
        assert abs(len(dh.df) - len(ref)) <= 0.02 * len(ref)  # points within the LAS quantum of a voxel face may change sides
        assert set(np.unique(dh.df["change_type"].to_numpy()).astype(int).tolist()) <= {1, 2, 3, 4}
