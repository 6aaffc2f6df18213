"""CPU tests of the neighbour-query code: the traversal, nearest-neighbour and union-find
routines of csrc/neighbors.cu are __host__ __device__; the library's vapc_selftest_*_host entry points run them
serially on host arrays.  Checked against the oracle (scipy KD-tree restatement of the reference).  No GPU."""
import ctypes as C

import numpy as np
import pytest

from oracle import vapc_oracle as O


def build_cell_list(x, y, z, h, origin=(0.0, 0.0, 0.0)):
    """numpy cell list in the layout a VoxelCloud of (x, y, z) with voxel size h has."""
    from vapc_b200 import _lib as L

    g = L.Grid()
    g.origin[:] = origin
    g.voxel_size = h
    mm = (C.c_double * 6)(x.min(), y.min(), z.min(), x.max(), y.max(), z.max())
    L.check(L.raw("vapc_grid_from_bounds")(C.byref(g), mm))
    cells = [np.floor((c - o) / h).astype(np.int64) - int(g.vmin[k]) for k, (c, o) in enumerate(zip((x, y, z), origin))]
    key = (cells[0].astype(np.uint64) << np.uint64(g.bits[1] + g.bits[2])) | (cells[1].astype(np.uint64) << np.uint64(g.bits[2])) | cells[2].astype(np.uint64)
    order = np.argsort(key, kind="stable")
    ks = key[order]
    head = np.ones(len(ks), bool)
    head[1:] = ks[1:] != ks[:-1]
    cell_keys = ks[head].astype(np.uint32 if g.key_bytes == 4 else np.uint64)
    start = np.append(np.flatnonzero(head), len(ks)).astype(np.uint32)
    rec = np.empty((len(x), 4), np.float64)
    rec[:, 0], rec[:, 1], rec[:, 2] = x, y, z
    rec[:, 3] = np.arange(len(x), dtype=np.int64).view(np.float64)
    return g, cell_keys, start, order.astype(np.uint32), rec


def P(a):
    return a.ctypes.data_as(C.c_void_p)


def host_nn(q, pts, h, bounded=False):
    from vapc_b200 import _lib as L

    g, keys, start, pos, rec = build_cell_list(pts[:, 0].copy(), pts[:, 1].copy(), pts[:, 2].copy(), h)
    qx, qy, qz = (np.ascontiguousarray(q[:, k]) for k in range(3))
    idx = np.empty(len(q), np.uint32)
    dist = np.empty(len(q), np.float64)
    L.call("vapc_selftest_cell_nn_host", P(qx), P(qy), P(qz), len(q), C.byref(g), P(keys), len(keys), P(start), P(pos), P(rec), len(pts), int(bounded),
           P(idx), P(dist))
    return idx.astype(np.int64), dist


def host_clusters(pts, d, first=None, weight=None):
    from vapc_b200 import _lib as L

    h = d * (1 + 2.0 ** -30) if d > 0 else 1.0
    g, keys, start, pos, rec = build_cell_list(pts[:, 0].copy(), pts[:, 1].copy(), pts[:, 2].copy(), h)
    n = len(pts)
    root, hop2, cfirst = (np.empty(n, np.uint32) for _ in range(3))
    csize = np.empty(n, np.int64)
    f = None if first is None else first.astype(np.uint32)
    w = None if weight is None else weight.astype(np.uint32)
    L.call("vapc_selftest_cluster_host", C.byref(g), P(keys), len(keys), P(start), P(pos), P(rec), n, float(d), None if f is None else P(f),
           None if w is None else P(w), P(root), P(hop2), P(cfirst), P(csize))
    return root.astype(np.int64), hop2.astype(np.int64), cfirst.astype(np.int64), csize


def labels_from(root, hop2, cfirst, csize, first=None):
    first = np.arange(len(root)) if first is None else first
    opens = np.sort(first[hop2 == first])
    return np.searchsorted(opens, cfirst[root]) + 1, csize[root]


@pytest.mark.parametrize("h", [0.05, 0.4, 3.0])
@pytest.mark.parametrize("seed", [0, 1])
def test_nearest_neighbour_equals_kdtree(seed, h):
    from scipy.spatial import KDTree

    rng = np.random.default_rng(seed)
    pts = rng.random((3000, 3)) * [20.0, 10.0, 2.0]
    pts[:400] = rng.normal([5, 5, 1], 0.05, (400, 3))  # a dense blob: many points per cell
    q = np.concatenate([rng.random((2000, 3)) * [20.0, 10.0, 2.0], rng.random((50, 3)) * 200 - 100, pts[:100]])  # inside, far outside, exact hits
    idx, dist = host_nn(q, pts, h)
    d_ref, i_ref = KDTree(pts).query(q, k=1)
    np.testing.assert_array_equal(idx, i_ref)
    np.testing.assert_array_equal(dist, d_ref)  # same formula: bit-exact


def test_nearest_neighbour_ties_take_lowest_index_and_single_point_sets():
    pts = np.array([[1.0, 0, 0], [-1.0, 0, 0], [0, 1.0, 0], [0, 0, 7.0]])
    idx, dist = host_nn(np.array([[0.0, 0, 0], [0, 0, 6.5]]), pts, 0.5)
    assert idx.tolist() == [0, 3] and dist.tolist() == [1.0, 0.5]
    idx, dist = host_nn(np.array([[3.0, 4.0, 0.0]]), np.array([[0.0, 0.0, 0.0]]), 1.0)
    assert idx.tolist() == [0] and dist.tolist() == [5.0]


def test_mutual_nearest_neighbours_equal_reference_statement():
    rng = np.random.default_rng(5)
    a = rng.random((1500, 3)) * 10
    b = np.concatenate([a[:900] + rng.normal(0, 0.02, (900, 3)), rng.random((500, 3)) * 10])
    i1, i2, d = O.mutual_nearest_neighbors(a, b)
    fwd, dist = host_nn(a, b, 0.5)
    back, _ = host_nn(b, a, 0.5)
    mutual = back[fwd] == np.arange(len(a))
    np.testing.assert_array_equal(np.flatnonzero(mutual), i1)
    np.testing.assert_array_equal(fwd[mutual], i2)
    np.testing.assert_array_equal(dist[mutual], d)


@pytest.mark.parametrize("case", ["grid_d1", "grid_d1.8", "float_d0.3", "blobs", "duplicates_d0"])
def test_clusters_equal_the_reference_sweep(case):
    rng = np.random.default_rng(11)
    if case.startswith("grid"):
        pts, d = rng.integers(0, 12, (300, 3)).astype(np.float64), float(case.split("_d")[1])
    elif case == "float_d0.3":
        pts, d = rng.random((500, 3)) * 5, 0.3
    elif case == "blobs":
        pts, d = np.concatenate([rng.normal(c, 0.3, (15, 3)) for c in ([0, 0, 0], [5, 0, 0], [0, 7, 1])]), 0.9
    else:
        pts, d = rng.integers(0, 4, (60, 3)).astype(np.float64), 0.0
    ref_id, ref_size = O.compute_clusters(pts, d)  # literal restatement of vapc.py:1336-1357
    cid, csz = labels_from(*host_clusters(pts, d))
    np.testing.assert_array_equal(cid, ref_id.astype(np.int64))
    np.testing.assert_array_equal(csz, ref_size.astype(np.int64))


def test_clusters_over_weighted_units_equal_clusters_over_their_points():
    """Occupied voxels as units (first = lowest point index, weight = point count) give the per-point result."""
    rng = np.random.default_rng(3)
    cells = rng.integers(0, 9, (40, 3))
    pts = cells[rng.integers(0, 40, 45)].astype(np.float64)  # few points per cell: every closed neighbourhood <= 50 rows
    uniq, first, inv, cnt = np.unique(pts, axis=0, return_index=True, return_inverse=True, return_counts=True)
    ref_id, ref_size = O.compute_clusters(pts, 1.0)
    root, hop2, cfirst, csize = host_clusters(uniq, 1.0, first=first, weight=cnt)
    cid, csz = labels_from(root, hop2, cfirst, csize, first=first)
    np.testing.assert_array_equal(cid[inv.ravel()], ref_id.astype(np.int64))
    np.testing.assert_array_equal(csz[inv.ravel()], ref_size.astype(np.int64))


def test_bounded_search_gives_up_only_on_far_neighbours():
    """bounded mode (what the GPU path runs first): near queries get the exact answer, queries whose neighbour is many cells away
    are handed back (index 0xFFFFFFFF) for the all-pairs kernel — never a wrong answer."""
    from scipy.spatial import KDTree

    rng = np.random.default_rng(8)
    pts = rng.random((20000, 3)) * [20.0, 10.0, 2.0]
    q = np.concatenate([rng.random((3000, 3)) * [20.0, 10.0, 2.0], rng.random((500, 3)) * [20.0, 10.0, 2.0] + [60.0, 0.0, 0.0]])
    h = 0.25
    idx, dist = host_nn(q, pts, h, bounded=True)
    d_ref, i_ref = KDTree(pts).query(q, k=1)
    given_up = idx == 0xFFFFFFFF
    assert given_up[3000:].all() and not given_up[:3000].any()  # 40 m = 160 cells away / inside the cloud
    np.testing.assert_array_equal(idx[~given_up], i_ref[~given_up])
    np.testing.assert_array_equal(dist[~given_up], d_ref[~given_up])
    assert np.isinf(dist[given_up]).all()
