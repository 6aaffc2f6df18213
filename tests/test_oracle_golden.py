"""CPU tests: the oracle (oracle/vapc_oracle.py) against the golden fixtures produced by the
unmodified reference (oracle/make_golden.py), and against the known answers of the
reference's own unit tests.  These pin the oracle before it is used to judge the CUDA path."""
import numpy as np
import pytest

from oracle import vapc_oracle as O
from tests import parity as P


def _table(gold):
    x, y, z = gold["in_X"], gold["in_Y"], gold["in_Z"]
    vs, origin = float(gold["in_voxel_size"]), gold["in_origin"].tolist()
    vx, vy, vz = O.voxelize(x, y, z, origin, vs)
    return x, y, z, vs, origin, vx, vy, vz, O.group_by_voxel(vx, vy, vz)


def test_kat_voxelize_bit_exact():
    g = P.load_golden("kat_voxelize")
    for tag in "abc":
        vx, vy, vz = O.voxelize(g["in_X"], g["in_Y"], g["in_Z"], g[f"{tag}_origin"].tolist(), float(g[f"{tag}_vs"]))
        for got, d in zip((vx, vy, vz), "xyz"):
            np.testing.assert_array_equal(got, g[f"{tag}_voxel_{d}"])
    # SURVEY appendix C KAT-1, literal values (a multiply-by-reciprocal would give 3, 6, 7)
    vx, _, _ = O.voxelize(np.array([0.3, 0.6, 0.7, -1e-17, 5429100.35]), np.zeros(5), np.zeros(5), [0, 0, 0], 0.1)
    assert vx.tolist() == [2, 5, 6, -1, 54291003]


@pytest.mark.parametrize("name", P.cloud_cases())
def test_integer_columns_bit_exact(name):
    gold = P.load_golden(name)
    x, y, z, vs, origin, vx, vy, vz, g = _table(gold)
    np.testing.assert_array_equal(vx, gold["pt_voxel_x"])
    np.testing.assert_array_equal(vy, gold["pt_voxel_y"])
    np.testing.assert_array_equal(vz, gold["pt_voxel_z"])
    np.testing.assert_array_equal(O.point_count(g)[g.point2voxel], gold["pt_point_count"])
    assert O.percentage_occupied(g) == float(gold["percentage_occupied"])


@pytest.mark.parametrize("name", P.cloud_cases())
def test_float_columns(name):
    gold = P.load_golden(name)
    x, y, z, vs, origin, vx, vy, vz, g = _table(gold)
    p2v, V = g.point2voxel, g.nvoxels
    P.assert_close_rel(O.point_density(g, vs)[p2v], gold["pt_point_density"], what="density")
    cog = O.center_of_gravity(x, y, z, g)
    std = O.std_of_cog(x, y, z, g)
    for k, d in enumerate("xyz"):
        P.assert_close_rel(cog[p2v, k], gold[f"pt_cog_{d}"], rtol=4 * P.EPS, what=f"cog_{d}")
        P.assert_close_rel(std[p2v, k], gold[f"pt_std_{d}"], what=f"std_{d}", atol=8 * P.EPS * max(np.abs(c).max() for c in (x, y, z)))
    ctr, cor = O.center_of_voxel(g, origin, vs), O.corner_of_voxel(g, origin, vs)
    for k, d in enumerate("xyz"):
        np.testing.assert_array_equal(ctr[p2v, k], gold[f"pt_center_{d}"])
        np.testing.assert_array_equal(cor[p2v, k], gold[f"pt_corner_{d}"])
    # covariance: the oracle's reference-formula variant must reproduce the reference closely
    # (same formula, sums differ only in the last bits, amplified by cancellation -> H1 bound)
    ref6 = np.stack([P.per_voxel(gold, p2v, V, c) for c in ("cov_xx", "cov_xy", "cov_xz", "cov_yy", "cov_yz", "cov_zz")], axis=1)
    ssq = (O._group_sum(x * x, g) + O._group_sum(y * y, g) + O._group_sum(z * z, g)) / g.counts
    P.assert_cov_h1(O.covariance_reference_formula(x, y, z, g), ref6, g.counts, ssq, what="cov_ref")
    P.assert_cov_h1(O.covariance_exact(x, y, z, g), ref6, g.counts, ssq, what="cov_exact")
    # ... and the restatement of pandas' Kahan group sums reproduces the reference's columns BIT FOR BIT, noise included
    if len(x) <= 60000:
        np.testing.assert_array_equal(O.covariance_reference_bitwise(x, y, z, g), ref6, err_msg="cov_** of the reference, bit for bit")
    for a, b in (("cov_xy", "cov_yx"), ("cov_xz", "cov_zx"), ("cov_yz", "cov_zy")):
        np.testing.assert_array_equal(gold["pt_" + a], gold["pt_" + b])
    # eigenvalues + features
    ref_eig = np.stack([P.per_voxel(gold, p2v, V, f"Eigenvalue_{i}") for i in (1, 2, 3)], axis=1)
    for i in (2, 3):  # dgeev's complex noise pairs (see parity.load_golden) are rounding noise
        im = gold.get(f"pt_Eigenvalue_{i}__imag")
        if im is not None:
            l1 = gold["pt_Eigenvalue_1"]
            assert np.all(np.abs(im)[~np.isnan(l1)] <= 1e-9 * l1[~np.isnan(l1)])
    for faithful in (True, False):
        eig = O.eigenvalues(x, y, z, g, faithful=faithful)
        P.assert_eig_h7(eig, ref_eig, g.counts, what=f"eig faithful={faithful}")
    feat, eig_n = O.geometric_features(O.eigenvalues(x, y, z, g, faithful=True))
    ref_feat = np.stack([P.per_voxel(gold, p2v, V, c) for c in P.FEATURES], axis=1)
    P.assert_features_h7(feat, ref_feat, ref_eig, g.counts)
    # distances
    dist = O.distance_to_center_of_gravity(x, y, z, g, cog)
    ref_d = gold["pt_distance"]
    assert np.all(np.abs(dist - ref_d) <= 1e-9 * np.abs(ref_d) + 8 * P.EPS * vs)
    dmin, winner, is_min = O.closest_to_center_of_gravity(dist, g)
    assert np.all(np.abs(dmin[p2v] - gold["pt_min_distance"]) <= 1e-9 * gold["pt_min_distance"] + 8 * P.EPS * vs)
    # H5: the oracle's winner lies in the reference's tied set
    ref_min = gold["pt_min_distance"][winner]
    assert np.all(ref_d[winner] <= ref_min * (1 + 1e-12))
    # ... and with the reference's own centroid double (pandas' Kahan group mean) nothing is left to tolerance: cog_*, distance,
    # min_distance and the set of rows tied at the minimum equal the golden columns BIT FOR BIT
    if len(x) <= 60000:
        cog_b = O.center_of_gravity_reference_bitwise(x, y, z, g)
        for k, c in enumerate(("cog_x", "cog_y", "cog_z")):
            np.testing.assert_array_equal(cog_b[p2v, k], gold["pt_" + c], err_msg=c)
        dist_b = O.distance_to_center_of_gravity(x, y, z, g, cog_b)
        np.testing.assert_array_equal(dist_b, ref_d, err_msg="distance")
        dmin_b, _, is_min_b = O.closest_to_center_of_gravity(dist_b, g)
        np.testing.assert_array_equal(dmin_b[p2v], gold["pt_min_distance"], err_msg="min_distance")
        np.testing.assert_array_equal(is_min_b, ref_d == gold["pt_min_distance"], err_msg="rows tied at the minimum")
        std_b = O.std_reference_bitwise(x, y, z, g)  # pandas' Welford recurrence: the reference's std_* to the bit
        for k, c in enumerate(("std_x", "std_y", "std_z")):
            np.testing.assert_array_equal(std_b[p2v, k], gold["pt_" + c], err_msg=c)


@pytest.mark.parametrize("name", P.cloud_cases())
@pytest.mark.parametrize("mode", ["center_of_voxel", "corner_of_voxel", "center_of_gravity", "closest_to_center_of_gravity"])
def test_reduce_to_voxels(name, mode):
    gold = P.load_golden(name)
    x, y, z, vs, origin, vx, vy, vz, g = _table(gold)
    xyz, src = O.reduce_to_voxels(x, y, z, g, mode, origin, vs)
    ref = np.stack([gold[f"red_{mode}_{c}"] for c in "XYZ"], axis=1)
    if mode in ("center_of_voxel", "corner_of_voxel"):
        np.testing.assert_array_equal(xyz, ref)  # exact coordinates AND exact row order
    elif mode == "center_of_gravity":
        assert xyz.shape == ref.shape
        np.testing.assert_allclose(xyz, ref, rtol=1e-12, atol=0)
    else:
        # every reference row is one of the input points; compare as sets of point indices,
        # allowing the H5 near-tie slack (rows the reference kept because of exact fp ties)
        ref_set = {tuple(r) for r in ref.tolist()}
        got_set = {tuple(r) for r in xyz.tolist()}
        sym = ref_set ^ got_set
        assert len(sym) <= max(2, int(0.02 * len(ref_set))), f"{len(sym)} rows differ"
    # non-XYZ columns of a reduced row are those of its source point
    for c in [k[len(f"red_{mode}_"):] for k in gold if k.startswith(f"red_{mode}_") and not k.endswith("__columns")]:
        if c in "XYZ" or f"in_{c}" not in gold or mode == "closest_to_center_of_gravity":
            continue
        np.testing.assert_array_equal(gold[f"in_{c}"][src], gold[f"red_{mode}_{c}"])


@pytest.mark.parametrize("name", [n for n in P.cloud_cases() if any(k.startswith("stat_") for k in P.load_golden(n))])
def test_attribute_statistics(name):
    gold = P.load_golden(name)
    x, y, z, vs, origin, vx, vy, vz, g = _table(gold)
    for key in [k for k in gold if k.startswith("stat_") and not k.startswith("stat_voxel_")]:
        col = key[len("stat_"):]
        attr = next(a for a in sorted((k[3:] for k in gold if k.startswith("in_")), key=len, reverse=True) if col.startswith(a + "_"))
        stat = col[len(attr) + 1:]
        got = O.attribute_statistic(gold["in_" + attr], g, stat)[g.point2voxel]
        # the oracle makes the reference's numpy calls on the reference's slices (vapc.py:393-406): every statistic is the golden column
        np.testing.assert_array_equal(got, gold[key], err_msg=col)


@pytest.mark.parametrize("name", [n for n in P.cloud_cases() if "mask_in_rows" in P.load_golden(n)])
def test_mask_and_buffer(name):
    gold = P.load_golden(name)
    x, y, z, vs, origin, vx, vy, vz, g = _table(gold)
    mv = O.voxelize(gold["in_mask_X"], gold["in_mask_Y"], gold["in_mask_Z"], origin, vs)
    member = O.mask_membership(vx, vy, vz, *mv)
    np.testing.assert_array_equal(np.flatnonzero(member), gold["mask_in_rows"])
    np.testing.assert_array_equal(np.flatnonzero(~member), gold["mask_out_rows"])
    buf = O.voxel_buffer(*mv, buffer_size=int(gold["in_mask_buffer_size"]))
    np.testing.assert_array_equal(buf, gold["mask_buffer_voxels"])  # same rows, same order
    member = O.mask_membership(vx, vy, vz, buf[:, 0], buf[:, 1], buf[:, 2])
    np.testing.assert_array_equal(np.flatnonzero(member), gold["mask_buffer_in_rows"])


@pytest.mark.parametrize("name", [n for n in P.cloud_cases() if "bi_merged_voxels" in P.load_golden(n)])
def test_bitemporal(name):
    gold = P.load_golden(name)
    vs, alpha = float(gold["in_bi_voxel_size"]), float(gold["in_bi_alpha"])
    tabs = []
    for k in "12":
        x, y, z = gold[f"in_e{k}_X"], gold[f"in_e{k}_Y"], gold[f"in_e{k}_Z"]
        g = O.group_by_voxel(*O.voxelize(x, y, z, [0, 0, 0], vs))
        cog = O.center_of_gravity(x, y, z, g)
        cov = O.cov6_to_cov9(O.covariance_reference_formula(x, y, z, g))
        # the reference re-voxelises the centroid rows (bi_vapc.py:80-81) after sorting them
        # by centroid floats (vapc.py:1209)
        order = np.lexsort((cog[:, 2], cog[:, 1], cog[:, 0]))
        cog, cov = cog[order], cov[order]
        vox = np.stack(O.voxelize(cog[:, 0], cog[:, 1], cog[:, 2], [0, 0, 0], vs), axis=1)
        tabs.append((vox, cog, cov))
    i1, i2, only1, only2 = O.join_epochs(tabs[0][0], tabs[1][0])
    np.testing.assert_array_equal(tabs[0][0][i1], gold["bi_merged_voxels"])
    cog1, cog2 = tabs[0][1][i1], tabs[1][1][i2]
    P.assert_close_rel(O.euclidean_distance(cog1, cog2), gold["bi_merged_distance"], rtol=1e-9, what="distance")
    res = O.mahalanobis_change(cog1, tabs[0][2][i1], cog2, tabs[1][2][i2], alpha=alpha)
    eye = np.eye(3) * 1e-10
    for k, cov in (("d1", tabs[1][2][i2]), ("d2", tabs[0][2][i1])):
        P.assert_mahalanobis(res[k], gold["bi_" + k], cov.reshape(-1, 3, 3) + eye, cov_rel=1e-12, what=k)
    ref_p = gold["bi_merged_p_value"]
    assert np.all((np.abs(res["p_value"] - ref_p) <= 1e-15 + 1e-6 * np.abs(ref_p)) | (np.isnan(ref_p) & np.isnan(res["p_value"])))
    # classification is exact away from the alpha threshold
    near = np.abs(ref_p - alpha) < 1e-6
    np.testing.assert_array_equal(res["significance"][~near], gold["bi_merged_mahalanobi_significance"][~near])
    np.testing.assert_array_equal(res["change_type"][~near], gold["bi_merged_change_type"][~near])
    # appear / disappear rows
    n_m = len(i1)
    exp_vox = gold["bi_export_voxels"]
    np.testing.assert_array_equal(exp_vox[n_m:n_m + len(only1)], tabs[0][0][only1])
    np.testing.assert_array_equal(exp_vox[n_m + len(only1):], tabs[1][0][only2])
    ct = gold["bi_export_change_type"]
    assert (ct[n_m:n_m + len(only1)] == 3).all() and (ct[n_m + len(only1):] == 4).all()


def test_reference_unit_test_known_answers():
    """Known answers of the reference's own unit tests, replayed on the oracle."""
    g = P.load_golden("plane500_vs1.0")
    # tests/test_compute.py:336-351  point_count > 3 at voxel 1 keeps exactly 442 points
    x, y, z, vs, origin, vx, vy, vz, grp = _table(g)
    assert int(O.filter_mask(O.point_count(grp)[grp.point2voxel], ">", 3).sum()) == 442 == int(g["filter_point_count_gt3_rows"])
    # tests/test_compute.py:202-207  reduction point [120, 200, 3]
    assert [int(x.min()), int(y.min()), int(z.min())] == [120, 200, 3] == g["reduction_point"].tolist()
    with pytest.raises(ValueError):
        O.filter_mask(np.arange(3), "leq", 3)  # tests/test_compute.py:312-321
    # tests/test_voxel.py:88-106, :217-229  known 8-voxel grid, exact order of centre reduce
    g8 = P.load_golden("grid8_vs1")
    x, y, z, vs, origin, vx, vy, vz, grp = _table(g8)
    vox = np.stack([grp.vx, grp.vy, grp.vz], axis=1)
    exp = np.array(sorted(zip([0, 1, 1, 3, 4, 5, 6, 9], [0, 1, 2, 3, 5, 2, 5, 7], [0, 1, 0, 0, 1, 0, 2, 1])))
    np.testing.assert_array_equal(vox, exp)
    xyz, _ = O.reduce_to_voxels(x, y, z, grp, "center_of_voxel", origin, vs)
    np.testing.assert_array_equal(xyz, exp + 0.5)
    # tests/test_voxel.py:144-162  27-neighbour buffer set
    from itertools import product
    buf = O.voxel_buffer(vx, vy, vz, 1)
    want = np.unique((exp[:, None] + np.array(list(product(range(-1, 2), repeat=3)))).reshape(-1, 3), axis=0)
    np.testing.assert_array_equal(np.unique(buf, axis=0), want)


@pytest.mark.parametrize("name", ["uniform5000_vs0.5", "plane500_vs1.0", "kat_cloud10"])
def test_reference_port_reproduces_reference(name):
    """oracle/ref_port.py (the timed CPU baseline) keeps the reference's pandas data flow; its columns must
    equal the reference's."""
    from oracle import ref_port as R

    gold = P.load_golden(name)
    df = R.full_attribute_set(gold["in_X"], gold["in_Y"], gold["in_Z"], float(gold["in_voxel_size"]), gold["in_origin"].tolist())
    for c in df.columns:
        got = df[c].to_numpy()
        got = got.real if got.dtype.kind == "c" else got
        np.testing.assert_allclose(got, gold["pt_" + c], rtol=1e-12, atol=1e-14, equal_nan=True, err_msg=c)


# ---- compute_clusters / merge_vapcs_with_closest_point ---------------------------------
CLUSTER_CASES = [("sparse_d1.0", 1.0, None), ("sparse_d1.8", 1.8, None), ("sparse_d2.0", 2.0, None), ("xyz_d0.3", 0.3, "XYZ")]


def cluster_rows(gold, case, by):
    x, y, z = (gold[f"in_sparse_{c}"] for c in "XYZ")
    if by == "XYZ":
        return np.stack([x, y, z], 1)
    return np.stack(O.voxelize(x, y, z, (0.0, 0.0, 0.0), float(gold["in_sparse_voxel_size"])), 1)


@pytest.mark.parametrize("case,d,by", CLUSTER_CASES)
def test_clusters(case, d, by):
    """The literal restatement of the sweep AND its closed form (components + opening rule) reproduce the reference."""
    gold = P.load_golden("clusters_nn")
    rows = cluster_rows(gold, case, by)
    for fn in (O.compute_clusters, O.cluster_components):
        cid, size = fn(rows, d)
        assert cid.dtype == gold[f"cl_{case}_cluster_id"].dtype and size.dtype == gold[f"cl_{case}_cluster_size"].dtype
        np.testing.assert_array_equal(cid, gold[f"cl_{case}_cluster_id"])
        np.testing.assert_array_equal(size, gold[f"cl_{case}_cluster_size"])


def test_clusters_reference_unit_test_fixture():
    """tests/test_compute.py:235-242: three blobs -> three clusters (every KD-tree query is truncated to 50 rows here)."""
    gold = P.load_golden("clusters_nn")
    rows = np.stack(O.voxelize(*(gold[f"in_blobs_{c}"] for c in "XYZ"), (0.0, 0.0, 0.0), 0.2), 1)
    cid, size = O.compute_clusters(rows, 10.0)
    np.testing.assert_array_equal(cid, gold["cl_blobs_cluster_id"])
    np.testing.assert_array_equal(size, gold["cl_blobs_cluster_size"])
    assert len(np.unique(cid)) == 3
    comp, _ = O.cluster_components(rows, 10.0)
    assert len(set(zip(comp.tolist(), cid.tolist()))) == 3  # same partition


def epoch_centroid_rows(gold, k):
    """reduce_to_voxels(center_of_gravity) of one epoch: centroid rows sorted by (X, Y, Z) (vapc.py:1150-1210)."""
    x, y, z = (gold[f"in_nn_e{k}_{c}"] for c in "XYZ")
    g = O.group_by_voxel(*O.voxelize(x, y, z, (0.0, 0.0, 0.0), float(gold["in_nn_voxel_size"])))
    cog = O.center_of_gravity(x, y, z, g)
    return cog[np.lexsort((cog[:, 2], cog[:, 1], cog[:, 0]))]


def test_mutual_nearest_neighbours():
    gold = P.load_golden("clusters_nn")
    c1, c2 = epoch_centroid_rows(gold, 1), epoch_centroid_rows(gold, 2)
    i1, i2, dist = O.mutual_nearest_neighbors(c1, c2)
    assert len(i1) == len(gold["nn_distance"])
    for k, c in enumerate("xyz"):
        P.assert_close_rel(c1[i1, k], gold[f"nn_cog_{c}_x"], rtol=1e-12, atol=1e-13, what=f"cog_{c}_x")
        P.assert_close_rel(c2[i2, k], gold[f"nn_cog_{c}_y"], rtol=1e-12, atol=1e-13, what=f"cog_{c}_y")
    P.assert_close_rel(dist, gold["nn_distance"], rtol=1e-9, atol=1e-13, what="distance")
