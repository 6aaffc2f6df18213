"""GPU tests of the neighbour queries: compute_clusters and merge_vapcs_with_closest_point of the
facade against the fixture the UNMODIFIED reference produced (tests/golden/clusters_nn.npz, oracle/make_golden.py),
and the kernels at larger sizes against the oracle (scipy KD-tree restatement).  Integer outputs (cluster ids, sizes,
neighbour indices) bit-exact; neighbour distances bit-exact against scipy on identical coordinates (same formula),
1e-9 where the coordinates are centroids computed on the GPU."""
import numpy as np
import pandas as pd
import pytest

from oracle import vapc_oracle as O
from tests import parity as P

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def vapc():
    import vapc_b200

    return vapc_b200


def sparse_vapc(vapc, gold):
    v = vapc.Vapc(float(gold["in_sparse_voxel_size"]))
    v.df = pd.DataFrame({c: gold[f"in_sparse_{c}"] for c in "XYZ"})
    v.original_attributes = v.df.columns.tolist()
    return v


@pytest.mark.parametrize("case,d,by", [("sparse_d1.0", 1.0, None), ("sparse_d1.8", 1.8, None), ("sparse_d2.0", 2.0, None),
                                       ("xyz_d0.3", 0.3, ["X", "Y", "Z"])])
def test_compute_clusters_matches_reference(vapc, case, d, by):
    gold = P.load_golden("clusters_nn")
    v = sparse_vapc(vapc, gold)
    v.compute_clusters(cluster_distance=d, cluster_by=by)
    assert v.df.columns.tolist() == gold[f"cl_{case}__columns"].tolist()
    for col in ("cluster_id", "cluster_size"):
        ref = gold[f"cl_{case}_{col}"]
        assert v.df[col].dtype == ref.dtype, col
        np.testing.assert_array_equal(v.df[col].to_numpy(), ref)
    assert isinstance(v.df.index, pd.RangeIndex)
    assert v.drop_columns[-3:] == ["voxel_x", "voxel_y", "voxel_z"]


def test_compute_clusters_reference_unit_test(vapc):
    """tests/test_compute.py:128-160, 235-242 of the reference: 3 blobs, voxel 0.2, distance 10 voxels -> 3 clusters.
    The reference's queries are truncated to 50 rows here; its partition and even its ids still come out equal."""
    gold = P.load_golden("clusters_nn")
    v = vapc.Vapc(0.2)
    v.df = pd.DataFrame({"X": gold["in_blobs_X"], "Y": gold["in_blobs_Y"], "Z": gold["in_blobs_Z"], "cluster_id": gold["in_blobs_true_id"]})
    v.original_attributes = v.df.columns.tolist()
    v.compute_clusters(cluster_distance=10.0)
    assert "cluster_id" in v.df.columns and "cluster_size" in v.df.columns
    assert v.df["cluster_id"].nunique() == 3
    np.testing.assert_array_equal(v.df["cluster_id"].to_numpy(), gold["cl_blobs_cluster_id"])
    np.testing.assert_array_equal(v.df["cluster_size"].to_numpy(), gold["cl_blobs_cluster_size"])
    with pytest.raises(KeyError):
        v.compute_clusters(cluster_distance=1.0, cluster_by=["nope", "Y", "Z"])


def test_merge_vapcs_with_closest_point_matches_reference(vapc):
    gold = P.load_golden("clusters_nn")
    eps = []
    for k in "12":
        v = vapc.Vapc(float(gold["in_nn_voxel_size"]))
        v.df = pd.DataFrame({c: gold[f"in_nn_e{k}_{c}"] for c in "XYZ"})
        v.original_attributes = v.df.columns.tolist()
        eps.append(v)
    bi = vapc.BiTemporalVapc(eps)
    bi.prepare_data_for_mahalanobis_distance()
    bi.merge_vapcs_with_closest_point()
    assert bi.distance_computed
    md = bi.merged_df
    assert md.columns.tolist() == gold["nn__columns"].tolist()[: len(md.columns)]
    assert len(md) == len(gold["nn_distance"])
    # the rows follow the epoch-1 centroid sort; a last-bit centroid difference may swap equal-X neighbours, so align
    # the rows by their (rounded) epoch-1 centroid
    def order(x, y, z):
        return np.lexsort((np.round(z, 7), np.round(y, 7), np.round(x, 7)))
    og = order(*(md[f"cog_{c}_x"].to_numpy() for c in "xyz"))
    orf = order(*(gold[f"nn_cog_{c}_x"] for c in "xyz"))
    assert (og == orf).mean() > 0.99
    for c in ("cog_x_x", "cog_y_x", "cog_z_x", "cog_x_y", "cog_y_y", "cog_z_y", "X_x", "X_y"):
        P.assert_close_rel(md[c].to_numpy()[og], gold["nn_" + c][orf], rtol=1e-9, atol=1e-12, what=c)
    P.assert_close_rel(md["distance"].to_numpy()[og], gold["nn_distance"][orf], rtol=1e-9, atol=1e-12, what="distance")
    bi.compute_mahalanobis_distance(alpha=0.005)
    assert bi.merged_df.columns.tolist() == gold["nn__columns"].tolist()
    p, pr = bi.merged_df["p_value"].to_numpy(), gold["nn_p_value"]
    p, pr = p[og], pr[orf]
    assert np.array_equal(np.isnan(p), np.isnan(pr))
    clear = ~np.isnan(pr) & (np.abs(pr - 0.005) > 1e-6)
    np.testing.assert_array_equal(bi.merged_df["mahalanobi_significance"].to_numpy()[og][clear], gold["nn_mahalanobi_significance"][orf][clear])


@pytest.mark.parametrize("n,nq,h", [(200_000, 300_000, 0.05), (200_000, 50_000, 1.5), (1_000, 20_000, 0.01)])
def test_nearest_neighbours_equal_kdtree_at_scale(n, nq, h):
    from scipy.spatial import KDTree

    from vapc_b200 import engine as E

    rng = np.random.default_rng(n + nq)
    pts = rng.random((n, 3)) * [30.0, 20.0, 3.0]
    pts[: n // 10] = rng.normal([10, 10, 1.5], 0.2, (n // 10, 3))  # dense blob
    q = np.concatenate([rng.random((nq - 200, 3)) * [30.0, 20.0, 3.0], rng.random((100, 3)) * 400 - 200, pts[:100]])
    cells = E._cell_cloud(pts[:, 0].copy(), pts[:, 1].copy(), pts[:, 2].copy(), h)
    idx, dist = E.nearest_neighbors(q[:, 0].copy(), q[:, 1].copy(), q[:, 2].copy(), cells)
    d_ref, i_ref = KDTree(pts).query(q, k=1, workers=-1)
    np.testing.assert_array_equal(idx.cpu().numpy(), i_ref)
    np.testing.assert_array_equal(dist.cpu().numpy(), d_ref)
    # mutual pairs == the reference's statement
    a, b = pts[:5000], q[:6000]
    i1, i2, d = E.mutual_nearest_neighbors([a[:, k].copy() for k in range(3)], [b[:, k].copy() for k in range(3)])
    r1, r2, rd = O.mutual_nearest_neighbors(a, b)
    np.testing.assert_array_equal(i1.cpu().numpy(), r1)
    np.testing.assert_array_equal(i2.cpu().numpy(), r2)
    np.testing.assert_array_equal(d.cpu().numpy(), rd)


def test_module_level_helpers_of_bi_vapc():
    """bi_vapc.mutual_nearest_neighbors / merge_with_suffix / compute_euclidean_distance (bi_vapc.py:171-207), which notebooks import
    by name: pairs equal the reference's KD-tree statement, the distance is the reference's formula within one rounding."""
    from vapc_b200 import bi_vapc as B

    rng = np.random.default_rng(77)
    a = rng.random((4000, 3)) * [30.0, 20.0, 3.0]
    b = a[rng.permutation(4000)[:3500]] + rng.normal(0, 0.05, (3500, 3))
    df1 = pd.DataFrame(a, columns=["cog_x", "cog_y", "cog_z"])
    df2 = pd.DataFrame(b, columns=["cog_x", "cog_y", "cog_z"])
    i1, i2 = B.mutual_nearest_neighbors(df1, df2)
    r1, r2, rd = O.mutual_nearest_neighbors(a, b)
    np.testing.assert_array_equal(i1, r1)
    np.testing.assert_array_equal(i2, r2)
    merged = B.merge_with_suffix(df1.iloc[i1].reset_index(drop=True), df2.iloc[i2].reset_index(drop=True))
    assert merged.columns.tolist() == ["cog_x_x", "cog_y_x", "cog_z_x", "cog_x_y", "cog_y_y", "cog_z_y"]
    dist = B.compute_euclidean_distance(merged)
    assert isinstance(dist, pd.Series) and dist.index.equals(merged.index)
    P.assert_close_rel(dist.to_numpy(), rd, rtol=1e-13, atol=0, what="distance")
    with pytest.raises(ValueError):
        B.mutual_nearest_neighbors(df1, df2, coord_cols=["cog_x", "cog_y"])


@pytest.mark.parametrize("kind,n,d", [("grid", 300_000, 1.0), ("grid", 300_000, 1.8), ("float", 200_000, 0.12), ("dups", 100_000, 0.0)])
def test_clusters_equal_components_at_scale(kind, n, d):
    """Against the closed form of the reference's sweep (oracle cluster_components, pinned to the reference by
    tests/test_oracle_golden.py); the Python sweep itself would take hours at this size."""
    from vapc_b200 import engine as E

    rng = np.random.default_rng(n)
    if kind == "grid":  # voxel indices of a sparse cloud incl. repeated rows (several points per voxel)
        rows = rng.integers(0, 90, (n, 3)).astype(np.int64)
    elif kind == "float":
        rows = rng.random((n, 3)) * [10.0, 10.0, 2.0]
    else:
        rows = rng.integers(0, 30, (n, 3)).astype(np.float64)  # float rows with exact duplicates, distance 0
    ref_id, ref_size = O.cluster_components(rows, d, dtype=np.int64)
    label, size, opens = E.cluster_components(rows[:, 0].copy(), rows[:, 1].copy(), rows[:, 2].copy(), d, integer_rows=(kind == "grid"))
    np.testing.assert_array_equal(label.cpu().numpy(), ref_id)
    np.testing.assert_array_equal(size.cpu().numpy(), ref_size.astype(np.int64))
    assert opens >= len(np.unique(ref_id))
    # run-to-run identical (the union-find is order-free: roots are component minima)
    label2, size2, _ = E.cluster_components(rows[:, 0].copy(), rows[:, 1].copy(), rows[:, 2].copy(), d, integer_rows=(kind == "grid"))
    assert bool((label == label2).all()) and bool((size == size2).all())


def test_nearest_neighbours_of_sets_far_apart():
    """Two clouds 500 m apart (2000 cells): the grid search gives every query up, the tiled all-pairs kernel answers — still the
    KD-tree's indices and distances, at a bounded cost (a pure grid search would scan the whole set per query)."""
    from scipy.spatial import KDTree

    from vapc_b200 import engine as E

    rng = np.random.default_rng(12)
    a = rng.random((30_000, 3)) * [20.0, 20.0, 3.0]
    b = rng.random((25_000, 3)) * [20.0, 20.0, 3.0] + [520.0, 0.0, 0.0]
    b[:50] = a[:50] + 0.01  # a few genuine neighbours inside the other cloud
    cells = E._cell_cloud(b[:, 0].copy(), b[:, 1].copy(), b[:, 2].copy(), 0.25)
    idx, dist = E.nearest_neighbors(a[:, 0].copy(), a[:, 1].copy(), a[:, 2].copy(), cells)
    d_ref, i_ref = KDTree(b).query(a, k=1, workers=-1)
    np.testing.assert_array_equal(idx.cpu().numpy(), i_ref)
    np.testing.assert_array_equal(dist.cpu().numpy(), d_ref)
