"""GPU parity tests of the CUDA path (through the C ABI, via vapc_b200.engine) against the CPU
oracle on seeded inputs, plus size-independent properties at larger sizes.  Tolerances are the
parity rules of tests/parity.py."""
import numpy as np
import pytest
import torch

from oracle import vapc_oracle as O
from tests import parity as P

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def E():
    from vapc_b200 import engine

    return engine


def _np(t):
    return t.detach().cpu().numpy()


def _u32(t):
    return _np(t).astype(np.int64) & 0xFFFFFFFF


def make_cloud(kind, n, seed):
    rng = np.random.default_rng(seed)
    if kind == "uniform":  # LAS-like quantised uniform box
        q = [rng.integers(0, hi, n) for hi in (200000, 150000, 30000)]
        pts = np.stack(q, axis=1) * 1e-4
    elif kind == "negative":  # negative coordinates, unquantised
        pts = rng.uniform([-37.5, -12.25, -3.0], [12.5, 40.0, 9.0], (n, 3))
    elif kind == "clustered":  # heavy-tailed voxel occupancy: tight blobs + sheet + sparse noise
        centers = rng.uniform(0, 40, (60, 3)) * [1, 1, 0.2]
        blob = centers[rng.integers(0, 60, n // 2)] + rng.normal(0, 0.05, (n // 2, 3))
        sheet = np.stack([rng.uniform(0, 40, n // 4), rng.uniform(0, 40, n // 4), np.zeros(n // 4)], axis=1)
        sheet[:, 2] = 0.3 * np.sin(sheet[:, 0] / 3) + rng.normal(0, 0.01, n // 4)
        noise = rng.uniform(0, 40, (n - n // 2 - n // 4, 3))
        pts = np.round(np.vstack([blob, sheet, noise]), 4)
        pts = pts[rng.permutation(n)]
    elif kind == "terrain":  # most voxels hold one or two points, the voxels of a thin sheet 20-60: the 8-lane tier of the reduction
        sheet = np.stack([rng.uniform(0, 30, n // 2), rng.uniform(0, 30, n // 2), np.zeros(n // 2)], axis=1)
        sheet[:, 2] = 2.0 + 0.2 * np.sin(sheet[:, 0] / 5) + rng.uniform(-0.03, 0.03, n // 2)
        sparse = rng.uniform([0, 0, 0], [30, 30, 25], (n - n // 2, 3))
        pts = np.round(np.vstack([sheet, sparse]), 4)
        pts = pts[rng.permutation(n)]
    elif kind == "utm":  # large coordinates: 64-bit keys when the origin stays at 0
        pts = rng.uniform(0, 25, (n, 3)) + [476850.0, 5429100.0, 312.0]
        pts = np.round(pts, 3)
    elif kind == "wide":  # a few dense patches spread over 2 km x 2 km x 50 m: a grid of ~40 key bits with populated voxels
        centers = rng.uniform([0, 0, 0], [2000, 2000, 50], (40, 3))
        pts = np.round(centers[rng.integers(0, 40, n)] + rng.uniform(-0.6, 0.6, (n, 3)), 4)
        pts[:2] = [[0.0, 0.0, 0.0], [2000.0, 2000.0, 50.0]]
    elif kind == "duplicates":  # many exactly identical points and n == 1 / n == 2 voxels
        base = rng.integers(0, 50, (n // 4, 3)) * 0.37
        pts = np.vstack([base, base, base[: n // 8], rng.uniform(0, 18.5, (n - 2 * (n // 4) - n // 8, 3))])
    else:
        raise ValueError(kind)
    return np.ascontiguousarray(pts[:, 0]), np.ascontiguousarray(pts[:, 1]), np.ascontiguousarray(pts[:, 2])


# --------------------------------------------------------------------------------------
def test_voxel_coords_bit_exact(E):
    g = P.load_golden("kat_voxelize")
    for tag in "abc":
        vs, origin = float(g[f"{tag}_vs"]), g[f"{tag}_origin"].tolist()
        got = E.voxel_coords(g["in_X"], g["in_Y"], g["in_Z"], vs, origin)
        for t, d in zip(got, "xyz"):
            np.testing.assert_array_equal(_np(t), g[f"{tag}_voxel_{d}"])
    rng = np.random.default_rng(1)
    for vs, origin in ((0.1, [0, 0, 0]), (0.3, [1.5, -2.25, 0.3]), (1.0, [0, 0, 0]), (0.05, [-7.0, 3.0, 0.0])):
        # values exactly on and next to voxel faces, where a reciprocal multiply goes wrong
        k = rng.integers(-5000, 5000, 20000)
        x = k * vs + origin[0]
        x = np.concatenate([x, np.nextafter(x, np.inf), np.nextafter(x, -np.inf), rng.uniform(-500, 500, 20000)])
        ref = O.voxelize(x, x[::-1].copy(), x * 0.5, origin, vs)
        got = E.voxel_coords(x, x[::-1].copy(), x * 0.5, vs, origin)
        for a, b in zip(got, ref):
            np.testing.assert_array_equal(_np(a), b)


@pytest.mark.parametrize("n", [1, 2, 31, 255, 4096, 4097, 12289, 1_000_003])
@pytest.mark.parametrize("bits,dtype", [(1, np.uint32), (7, np.uint32), (8, np.uint32), (17, np.uint32), (32, np.uint32),
                                        (33, np.uint64), (47, np.uint64), (64, np.uint64)])
def test_sort_pairs_stable(E, n, bits, dtype):
    rng = np.random.default_rng(n * 131 + bits)
    hi = (1 << bits) - 1
    keys = rng.integers(0, hi, n, dtype=np.uint64, endpoint=True).astype(dtype)
    if n > 100:  # skew: half of the keys share one digit pattern
        keys[: n // 2] &= dtype(hi & 0xFF00FF00FF00FF00)
    tk = torch.from_numpy(keys.view(np.int32 if dtype == np.uint32 else np.int64)).cuda()
    ks, idx = E.sort_pairs(tk, bits)
    order = np.argsort(keys, kind="stable")
    got_keys = _np(ks).view(dtype)
    np.testing.assert_array_equal(got_keys, keys[order])
    np.testing.assert_array_equal(_u32(idx), order)


def test_sort_pairs_wide_lookback_words(E):
    """n >= 2^30 elements per GPU switch the look-back words to 64 bits; force that path at a small n."""
    from vapc_b200 import _lib as L

    rng = np.random.default_rng(11)
    n = 300_001
    keys = rng.integers(0, 1 << 19, n, dtype=np.uint64).astype(np.uint32)
    L.call("vapc_sort_debug_wide_lookback", 1)
    try:
        ks, idx = E.sort_pairs(torch.from_numpy(keys.view(np.int32)).cuda(), 19)
        torch.cuda.synchronize()
    finally:
        L.call("vapc_sort_debug_wide_lookback", 0)
    order = np.argsort(keys, kind="stable")
    np.testing.assert_array_equal(_np(ks).view(np.uint32), keys[order])
    np.testing.assert_array_equal(_u32(idx), order)


@pytest.mark.parametrize("n,nseg,bits,dtype", [(1, 1, 5, np.uint32), (5000, 3, 9, np.uint32), (300_000, 256, 16, np.uint32),
                                                (300_000, 200, 24, np.uint64), (1_000_003, 37, 11, np.uint32), (70_000, 256, 40, np.uint64)])
def test_sort_pairs_segmented(E, n, nseg, bits, dtype):
    """Every segment sorted independently and stably by the low bits; empty, tiny and multi-tile segments mixed."""
    rng = np.random.default_rng(n + nseg)
    cuts = np.sort(rng.integers(0, n + 1, nseg - 1)) if nseg > 1 else np.zeros(0, np.int64)
    if nseg > 4:
        cuts[1] = cuts[0]  # an empty segment
    seg_start = np.concatenate([[0], cuts, [n]]).astype(np.int64)
    keys = rng.integers(0, (1 << bits) - 1, n, dtype=np.uint64, endpoint=True).astype(dtype)
    seg_of = np.searchsorted(seg_start, np.arange(n), side="right") - 1
    order = np.lexsort((np.arange(n), keys, seg_of))  # by segment, then key, then position
    tk = torch.from_numpy(keys.view(np.int32 if dtype == np.uint32 else np.int64)).cuda()
    ss = torch.from_numpy(seg_start.astype(np.uint32).view(np.int32)).cuda()
    ks, idx = E.sort_pairs_segmented(tk, bits, ss, nseg)
    np.testing.assert_array_equal(_np(ks).view(dtype), keys[order])
    np.testing.assert_array_equal(_u32(idx), order)


def test_sort_pairs_bit_range(E):
    """Only the bits [begin_bit, end_bit) are sorted (stable in everything else): one pass over the leading byte of an index."""
    rng = np.random.default_rng(9)
    for n, begin, end, dtype in ((200_000, 10, 18, np.uint32), (50_000, 3, 27, np.uint32), (70_000, 40, 61, np.uint64)):
        keys = rng.integers(0, (1 << end) - 1, n, dtype=np.uint64, endpoint=True).astype(dtype)
        payload = rng.permutation(n).astype(np.uint32)
        tk = torch.from_numpy(keys.view(np.int32 if dtype == np.uint32 else np.int64)).cuda()
        ks, idx = E.sort_pairs(tk, end, idx=torch.from_numpy(payload.view(np.int32)).cuda(), begin_bit=begin, preserve_keys=True)
        field = (keys.astype(np.uint64) >> np.uint64(begin)) & np.uint64((1 << (end - begin)) - 1)
        order = np.argsort(field, kind="stable")
        np.testing.assert_array_equal(_np(ks).view(dtype), keys[order])
        np.testing.assert_array_equal(_u32(idx), payload[order])
        np.testing.assert_array_equal(_np(tk).view(dtype), keys)  # preserved


def test_sort_pairs_with_payload_and_partial_bits(E):
    rng = np.random.default_rng(5)
    n = 50000
    keys = rng.integers(0, 1 << 20, n, dtype=np.uint64).astype(np.uint32)
    payload = rng.permutation(n).astype(np.uint32)
    ks, idx = E.sort_pairs(torch.from_numpy(keys.view(np.int32)).cuda(), 12, idx=torch.from_numpy(payload.view(np.int32)).cuda())
    order = np.argsort(keys & 0xFFF, kind="stable")  # only the low 12 bits are sorted
    np.testing.assert_array_equal(_np(ks).view(np.uint32), keys[order])
    np.testing.assert_array_equal(_u32(idx), payload[order])


# --------------------------------------------------------------------------------------
def check_cloud(E, x, y, z, vs, origin, eig_faithful=False, bucketed=True):
    cloud = E.VoxelCloud.build(x, y, z, vs, origin, bucketed=bucketed)
    vx, vy, vz = O.voxelize(x, y, z, origin, vs)
    g = O.group_by_voxel(vx, vy, vz)
    V = g.nvoxels
    assert cloud.nvoxels == V
    # integer outputs: bit-exact
    cvx, cvy, cvz = (_np(t) for t in cloud.voxel_coords())
    np.testing.assert_array_equal(cvx, g.vx)
    np.testing.assert_array_equal(cvy, g.vy)
    np.testing.assert_array_equal(cvz, g.vz)
    np.testing.assert_array_equal(_u32(cloud.count), g.counts)
    np.testing.assert_array_equal(_u32(cloud.first_idx), g.first_index)
    np.testing.assert_array_equal(_u32(cloud.point2voxel), g.point2voxel)
    np.testing.assert_array_equal(_u32(cloud.idx_sorted), g.order)  # stable: original order inside a voxel
    np.testing.assert_array_equal(_u32(cloud.start), np.append(g.starts, len(x)))
    assert cloud.percentage_occupied() == O.percentage_occupied(g)
    # floating point outputs
    st = cloud.stats(density=True, cog=True, std=True, cov=True, eig=True, eig_n=True, feat=True)
    np.testing.assert_array_equal(_np(st.density), O.point_density(g, vs))  # count / vs ** 3 with the cube as Python's pow forms it
    cog_ref = O.center_of_gravity(x, y, z, g)
    P.assert_close_rel(_np(st.cog).T, cog_ref, what="cog")
    std_ref = O.std_of_cog(x, y, z, g)
    P.assert_close_rel(_np(st.std).T, std_ref, what="std", atol=1e-12 * vs)
    cov_exact = O.covariance_exact(x, y, z, g)
    ssq = np.zeros(V)  # vs the exact oracle no allowance for the raw-moment formula is needed
    P.assert_cov_h1(_np(st.cov).T, cov_exact, g.counts, ssq, what="cov vs exact")
    cov_ref = O.covariance_reference_formula(x, y, z, g)
    ssq = (O._group_sum(x * x, g) + O._group_sum(y * y, g) + O._group_sum(z * z, g)) / g.counts
    P.assert_cov_h1(_np(st.cov).T, cov_ref, g.counts, ssq, what="cov vs reference formula")
    eig_ref = O.eigenvalues(x, y, z, g, faithful=eig_faithful)
    floor = P.noise_floor(max(np.abs(c).max() for c in (x, y, z)))
    P.assert_eig_h7(_np(st.eig).T, eig_ref, g.counts, abs_floor=floor)
    feat_ref, eign_ref = O.geometric_features(eig_ref)
    P.assert_features_h7(_np(st.feat).T, feat_ref, eig_ref, g.counts, abs_floor=floor)
    ok = (g.counts > 1) & (eig_ref[:, 0] > 1e9 * floor)
    assert np.all(np.abs(_np(st.eig_n).T[ok] - eign_ref[ok]) <= 1e-9)
    # centre / corner: bit-exact
    center, corner = cloud.center_corner()
    np.testing.assert_array_equal(_np(center).T, O.center_of_voxel(g, origin, vs))
    np.testing.assert_array_equal(_np(corner).T, O.corner_of_voxel(g, origin, vs))
    # closest to centroid (H5)
    dist_ref = O.distance_to_center_of_gravity(x, y, z, g, cog_ref)
    dmin_ref, winner_ref, _ = O.closest_to_center_of_gravity(dist_ref, g)
    md, am = cloud.closest_to_cog()
    am = _u32(am)
    scale = 8 * P.EPS * max(vs, np.abs(cog_ref).max())
    assert np.all(np.abs(_np(md) - dmin_ref) <= 1e-9 * dmin_ref + scale)
    assert np.all(g.point2voxel[am] == np.arange(V)), "winner must belong to its voxel"
    assert np.all(dist_ref[am] <= dmin_ref * (1 + 1e-12) + scale), "winner must be in the reference's tied set"
    # voxels with a clear winner (runner-up separated from the minimum): exact match required
    ds = dist_ref[g.order]
    vid_sorted = g.point2voxel[g.order]
    ds_in_voxel = ds[np.lexsort((ds, vid_sorted))]
    second = np.full(V, np.inf)
    multi = g.counts > 1
    second[multi] = ds_in_voxel[g.starts[multi] + 1]
    clear = second > dmin_ref * (1 + 1e-9) + 1e-12
    np.testing.assert_array_equal(am[clear], winner_ref[clear])
    pd_ = _np(cloud.point_distance())
    assert np.all(np.abs(pd_ - dist_ref) <= 1e-9 * dist_ref + scale)
    return cloud, g


@pytest.mark.parametrize("kind,n,vs,origin", [
    ("uniform", 60000, 0.5, [0, 0, 0]),
    ("uniform", 200000, 0.1, [0, 0, 0]),
    ("uniform", 150000, 0.02, [0, 0, 0]),       # 28 key bits: the largest grids that take the rank-bitmap point->voxel path
    ("uniform", 150000, 0.01, [0, 0, 0]),       # 31 key bits in 32-bit keys: point->voxel by grouped scatter
    ("negative", 50000, 0.3, [1.5, -2.25, 0.3]),
    ("clustered", 120000, 0.25, [0, 0, 0]),
    ("clustered", 120000, 5.0, [0, 0, 0]),      # voxels far larger than a 1024-point tile: carry chains
    ("uniform", 70000, 100.0, [0, 0, 0]),       # every point in one voxel
    ("utm", 40000, 0.5, [0, 0, 0]),             # 64-bit keys
    ("utm", 40000, 0.05, [476850.0, 5429100.0, 312.0]),
    ("duplicates", 40000, 0.37, [0, 0, 0]),
    ("uniform", 1, 1.0, [0, 0, 0]),
    ("uniform", 3, 1000.0, [0, 0, 0]),
    ("wide", 50000, 0.05, [0, 0, 0]),           # 2 km extent at 5 cm: 40 key bits -> 32-bit low keys inside the buckets
    ("terrain", 400000, 0.25, [0, 0, 0]),       # < 8 points per voxel on average, sheet voxels of 20-60: groups of 8 lanes per voxel
    ("uniform", 300000, 0.5, [0, 0, 0]),        # ~41 points per voxel: four lanes per voxel
])
def test_voxel_table_vs_oracle(E, kind, n, vs, origin):
    x, y, z = make_cloud(kind, n, seed=n + int(vs * 100))
    if kind == "negative":  # the path with records in input order + full sort stays covered
        check_cloud(E, x, y, z, vs, origin, bucketed=False)
    check_cloud(E, x, y, z, vs, origin)


def test_mode_count_vs_oracle(E):
    x, y, z = make_cloud("clustered", 80000, 3)
    rng = np.random.default_rng(4)
    attr = rng.integers(0, 9, len(x)).astype(np.float64)
    attr[rng.random(len(x)) < 0.5] = 2.0
    for vs in (0.05, 0.2, 1.0, 50.0):
        cloud = E.VoxelCloud.build(x, y, z, vs, [0, 0, 0])
        g = O.group_by_voxel(*O.voxelize(x, y, z, [0, 0, 0], vs))
        mode, counts = cloud.mode_count(attr, [0.1, 0.34, 0.5, 1.0])
        np.testing.assert_array_equal(_np(mode), O.attribute_statistic(attr, g, "mode"))
        for p, t in counts.items():
            np.testing.assert_array_equal(_np(t), O.attribute_statistic(attr, g, f"mode_count,{p}"), err_msg=f"vs={vs} p={p}")
    with pytest.raises(ValueError):
        cloud.mode_count(attr - 3.0, [0.1])


def test_mode_falls_back_to_histogram_mode_per_voxel(E):
    """vapc.py:407-412: a voxel whose values np.bincount rejects (negative scan angles) gets the Freedman-Diaconis histogram mode
    (utilities.py:114-159) — that voxel only; the others keep argmax(bincount).  The column then holds floats."""
    from vapc_b200.utilities import compute_mode_continuous

    rng = np.random.default_rng(12)
    n = 6000
    pts = rng.uniform(0, 3, (n, 3))
    x, y, z = (np.ascontiguousarray(pts[:, k]) for k in range(3))
    attr = rng.integers(0, 6, n).astype(np.float64)
    neg = pts[:, 0] < 1.0  # the voxels of the first metre hold negative, spread-out values
    attr[neg] = np.round(rng.normal(-5.0, 4.0, neg.sum()), 1)
    cloud = E.VoxelCloud.build(x, y, z, 1.0, [0, 0, 0])
    g = O.group_by_voxel(*O.voxelize(x, y, z, [0, 0, 0], 1.0))
    mode, _ = cloud.mode_count(attr, [], want_mode=True)
    assert mode.dtype == torch.float64
    got = _np(mode)
    for v in range(g.nvoxels):
        vals = attr[g.order[g.starts[v]:g.starts[v] + g.counts[v]]]
        if (vals < 0).any():
            assert got[v] == compute_mode_continuous(vals)
        else:
            assert got[v] == np.argmax(np.bincount(vals.astype(int)))
    with pytest.raises(ValueError):
        cloud.mode_count(attr, [0.1])  # mode_count has no fallback in the reference (vapc.py:423-428)
    # a flagged voxel whose histogram the reference cannot build (all values equal: zero bin width) raises what numpy raises there
    v0 = g.order[g.starts[0]:g.starts[0] + g.counts[0]]
    flat = attr.copy()
    flat[v0] = -2.0
    with pytest.raises(ValueError):
        compute_mode_continuous(flat[v0])
    with pytest.raises(ValueError):
        cloud.mode_count(flat, [], want_mode=True)


def test_attribute_statistics_vs_oracle(E):
    x, y, z = make_cloud("uniform", 30000, 6)
    rng = np.random.default_rng(7)
    attr = rng.normal(-3, 5, len(x))
    cloud = E.VoxelCloud.build(x, y, z, 0.7, [0, 0, 0])
    g = O.group_by_voxel(*O.voxelize(x, y, z, [0, 0, 0], 0.7))
    res = E.attribute_statistics(cloud, attr)
    for stat in ("min", "max", "median"):
        np.testing.assert_array_equal(_np(res[stat]), O.attribute_statistic(attr, g, stat), err_msg=stat)
    for stat in ("mean", "sum", "var", "std"):
        np.testing.assert_allclose(_np(res[stat]), O.attribute_statistic(attr, g, stat), rtol=1e-9, atol=1e-12, err_msg=stat)
    # the kernel adds the way numpy adds a slice (8 interleaved accumulators, pairwise halves beyond 128 values): sum / mean / var /
    # std of the reference's per-voxel numpy calls BIT FOR BIT — voxels of ~4, ~40, ~700 and all 30 000 values
    for vs in (0.7, 1.5, 4.0, 100.0):
        cloud = E.VoxelCloud.build(x, y, z, vs, [0, 0, 0])
        g = O.group_by_voxel(*O.voxelize(x, y, z, [0, 0, 0], vs))
        for scale in (1.0, 1e5):
            res = E.attribute_statistics(cloud, attr * scale + (scale - 1.0))
            for stat in ("mean", "sum", "var", "std", "min", "max", "median"):
                np.testing.assert_array_equal(_np(res[stat]), O.attribute_statistic(attr * scale + (scale - 1.0), g, stat), err_msg=f"{stat} vs={vs}")


def test_mask_and_buffer_vs_oracle(E):
    x, y, z = make_cloud("negative", 60000, 8)
    mx, my, mz = make_cloud("negative", 300, 9)
    vs, origin = 0.4, [0.1, 0.2, 0.3]
    vx, vy, vz = O.voxelize(x, y, z, origin, vs)
    mv = O.voxelize(mx, my, mz, origin, vs)
    mask = E.VoxelMask(*mv, vs, origin)
    ref = O.mask_membership(vx, vy, vz, *mv)
    np.testing.assert_array_equal(_np(mask.contains_points(x, y, z)).astype(bool), ref)
    np.testing.assert_array_equal(_np(mask.contains_voxels(vx, vy, vz)).astype(bool), ref)
    for b in (1, 2):
        want = O.voxel_buffer(*mv, buffer_size=b)
        uniq = E.unique_voxels_first_occurrence(*mv)
        exp = E.buffer_expand(uniq[0], uniq[1], uniq[2], b)
        got = E.unique_voxels_first_occurrence(exp[0], exp[1], exp[2])
        np.testing.assert_array_equal(np.stack([_np(t) for t in got], axis=1), want)  # same rows, same order
        bm = E.VoxelMask(want[:, 0], want[:, 1], want[:, 2], vs, origin)
        np.testing.assert_array_equal(_np(bm.contains_points(x, y, z)).astype(bool),
                                      O.mask_membership(vx, vy, vz, want[:, 0], want[:, 1], want[:, 2]))
    empty = E.VoxelMask(np.zeros(0, np.int64), np.zeros(0, np.int64), np.zeros(0, np.int64), vs, origin)
    assert not _np(empty.contains_points(x, y, z)).any()


def test_two_epoch_join_and_mahalanobis_vs_oracle(E):
    x1, y1, z1 = make_cloud("clustered", 60000, 10)
    rng = np.random.default_rng(11)
    keep = rng.random(len(x1)) < 0.85
    x2, y2, z2 = x1[keep] + rng.normal(0, 0.01, keep.sum()), y1[keep] + 0.02, z1[keep] + rng.normal(0, 0.01, keep.sum())
    ex, ey, ez = make_cloud("uniform", 3000, 12)
    x2, y2, z2 = np.concatenate([x2, ex + 41]), np.concatenate([y2, ey]), np.concatenate([z2, ez])
    vs = 0.5
    c1 = E.VoxelCloud.build(x1, y1, z1, vs, [0, 0, 0])
    c2 = E.VoxelCloud.build(x2, y2, z2, vs, [0, 0, 0])
    g1 = O.group_by_voxel(*O.voxelize(x1, y1, z1, [0, 0, 0], vs))
    g2 = O.group_by_voxel(*O.voxelize(x2, y2, z2, [0, 0, 0], vs))
    i1, i2, only1, only2 = E.join_voxel_tables(c1.voxel_coords(), c2.voxel_coords())
    r1, r2, ro1, ro2 = O.join_epochs(np.stack([g1.vx, g1.vy, g1.vz], 1), np.stack([g2.vx, g2.vy, g2.vz], 1))
    np.testing.assert_array_equal(_np(i1), r1)
    np.testing.assert_array_equal(_np(i2), r2)
    np.testing.assert_array_equal(_np(only1), ro1)
    np.testing.assert_array_equal(_np(only2), ro2)
    s1, s2 = c1.stats(cog=True, cov=True), c2.stats(cog=True, cov=True)
    for alpha in (0.005, 0.999):
        res = E.mahalanobis(s1.cog, s1.cov, s2.cog, s2.cov, i1, i2, alpha=alpha)
        cog1, cog2 = O.center_of_gravity(x1, y1, z1, g1)[r1], O.center_of_gravity(x2, y2, z2, g2)[r2]
        cv1 = O.cov6_to_cov9(O.covariance_exact(x1, y1, z1, g1))[r1]
        cv2 = O.cov6_to_cov9(O.covariance_exact(x2, y2, z2, g2))[r2]
        ref = O.mahalanobis_change(cog1, cv1, cog2, cv2, alpha=alpha)
        P.assert_close_rel(_np(res["distance"]), O.euclidean_distance(cog1, cog2), rtol=1e-9, atol=1e-13, what="distance")
        eye = np.eye(3) * 1e-10
        c_d1 = P.assert_mahalanobis(_np(res["d1"]), ref["d1"], cv2.reshape(-1, 3, 3) + eye, cov_rel=1e-9, what="d1")
        c_d2 = P.assert_mahalanobis(_np(res["d2"]), ref["d2"], cv1.reshape(-1, 3, 3) + eye, cov_rel=1e-9, what="d2")
        cond = np.full(len(r1), np.inf)
        cond[~np.isnan(ref["d1"])] = c_d1
        cond2 = np.full(len(r1), np.inf)
        cond2[~np.isnan(ref["d2"])] = c_d2
        well = (cond < 1e2) & (cond2 < 1e2)  # well-conditioned voxels: plain 1e-9-level agreement
        p, pr = _np(res["p_value"]), ref["p_value"]
        assert np.all(np.isnan(p) == np.isnan(pr))
        assert np.all(np.abs(p[well] - pr[well]) <= 1e-15 + 1e-9 * np.abs(pr[well])), float(np.max(np.abs(p[well] - pr[well]) / (1e-15 + 1e-9 * np.abs(pr[well]))))
        stable = well & (np.abs(pr - alpha) > 1e-6)
        np.testing.assert_array_equal(_np(res["significance"])[stable], ref["significance"][stable])
        np.testing.assert_array_equal(_np(res["change_type"])[stable], ref["change_type"][stable])
        nan_rows = np.isnan(pr)
        assert np.all(_np(res["change_type"])[nan_rows] == 0)


def test_lexsort_unique_rows_matches_pandas(E):
    """The GPU replacement of the reference's final drop_duplicates + groupby(XYZ) formatting visits the same rows in the
    same order as pandas (duplicates, -0.0 / 0.0, NaN coordinates, ties in X and Y)."""
    import pandas as pd

    rng = np.random.default_rng(3)
    for n in (1, 7, 5000, 200_000):
        x = rng.integers(-20, 20, n) * 0.25
        y = rng.integers(-5, 5, n) * 0.5
        z = np.round(rng.normal(0, 1, n), 1)
        if n > 100:
            x[::97] = -0.0
            x[1::97] = 0.0
            z[5::211] = np.nan
            y[7::313] = np.nan
        df = pd.DataFrame({"X": x, "Y": y, "Z": z, "row": np.arange(n, dtype=np.float64)})
        ref = df.drop_duplicates(subset=["X", "Y", "Z"]).groupby(["X", "Y", "Z"], as_index=False).median(numeric_only=True)
        rows = E.lexsort_unique_rows(x, y, z)
        np.testing.assert_array_equal(rows, ref["row"].to_numpy().astype(np.int64))


def test_plan_exchange_kernel_matches_host_mirror(E):
    """Multi-GPU exchange planning kernel (splitters + send-count matrix from the all-gathered histograms) == its host mirror
    (which tests/test_dist_cpu.py checks against the pure-Python planner)."""
    from vapc_b200 import _lib as L

    rng = np.random.default_rng(8)
    for world, hb, total_bits in ((1, 16, 24), (2, 16, 24), (3, 10, 10), (8, 16, 27), (8, 4, 4), (64, 16, 40), (5, 12, 63)):
        nb = 1 << hb
        hist = rng.integers(0, 1000, (world, nb)).astype(np.uint32)
        hist[rng.random((world, nb)) < 0.5] = 0
        if world > 2:
            hist[1] = 0
            hist[2, : nb // 2] = 0  # a rank whose rows all lie in the upper half of the key space
        ref = np.zeros(world * world + world - 1, dtype=np.int64)
        assert L.raw("vapc_selftest_plan_exchange_host")(hist.ctypes.data, world, nb, max(total_bits - hb, 0), ref.ctypes.data) == 0
        out = torch.empty(world * world + world - 1, dtype=torch.int64, device="cuda")
        L.call("vapc_plan_exchange", E._ptr(torch.from_numpy(hist.view(np.int32)).cuda()), world, hb, total_bits, E._ptr(out), E._stream())
        np.testing.assert_array_equal(out.cpu().numpy(), ref)
    # merge keys: own slice + first word of every received row
    gk = torch.arange(100, 200, dtype=torch.int64, device="cuda")
    rows = torch.zeros((7, 11), dtype=torch.float64, device="cuda")
    rows.view(torch.int64)[:, 0] = torch.arange(7, device="cuda") + (1 << 33)
    for kb, dt in ((4, torch.int32), (8, torch.int64)):
        out = torch.empty(30 + 7, dtype=dt, device="cuda")
        L.call("vapc_merge_keys", E._ptr(gk), 10, 30, E._ptr(rows), 7, kb, E._ptr(out), E._stream())
        expect = torch.cat([gk[10:40], torch.arange(7, device="cuda") + (1 << 33)])
        assert torch.equal(out.to(torch.int64) & (0xFFFFFFFF if kb == 4 else -1), expect & (0xFFFFFFFF if kb == 4 else -1))


@pytest.mark.parametrize("name", ["utm_offset_vs0.25", "uniform5000_vs0.25", "plane500_vs1.0", "vapc_in_every50th_vs0.25", "shifted_origin_vs0.3"])
def test_cog_reference_formula_bit_for_bit(E, name):
    """vapc_cog_reference_formula == the cog_* columns of the unmodified reference, bit for bit; with it as the cloud's centroid,
    vapc_point_distance == `distance`, vapc_closest_to_cog == `min_distance` and the winner is the FIRST row of the reference's exactly
    tied set — all assert_array_equal.  On a larger seeded cloud with duplicated points: == the oracle's Kahan restatement."""
    gold = P.load_golden(name)
    x, y, z = (np.asarray(gold[k], np.float64) for k in ("in_X", "in_Y", "in_Z"))
    vs, origin = float(gold["in_voxel_size"]), np.asarray(gold["in_origin"], float).tolist()
    for bucketed in (True, False):
        cloud = E.VoxelCloud.build(x, y, z, vs, origin, bucketed=bucketed)
        exact = _np(cloud.stats(cog=True).cog)
        cloud.centroid_formula = "reference"
        p2v = _u32(cloud.point2voxel)
        got = _np(cloud.stats(cog=True).cog)
        for k, c in enumerate(("cog_x", "cog_y", "cog_z")):
            np.testing.assert_array_equal(got[k][p2v], gold["pt_" + c], err_msg=f"{name} {c} bucketed={bucketed}")
        assert np.all(np.abs(got - exact) <= 8 * P.EPS * np.maximum(np.abs(got), vs))  # the two centroids differ in the last bits only
        np.testing.assert_array_equal(_np(cloud.point_distance()), gold["pt_distance"])
        std = _np(cloud.std_reference_formula())
        for k, c in enumerate(("std_x", "std_y", "std_z")):  # pandas' Welford recurrence: std_* of the reference to the bit
            np.testing.assert_array_equal(std[k][p2v], gold["pt_" + c], err_msg=f"{name} {c} bucketed={bucketed}")
        md, am = cloud.closest_to_cog()
        np.testing.assert_array_equal(_np(md)[p2v], gold["pt_min_distance"])
        tied = gold["pt_distance"] == gold["pt_min_distance"]
        first = np.full(cloud.nvoxels, len(x), np.int64)
        np.minimum.at(first, p2v[tied], np.nonzero(tied)[0])
        np.testing.assert_array_equal(_u32(am), first)
        cloud.centroid_formula = "exact"
        np.testing.assert_array_equal(_np(cloud.stats(cog=True).cog), exact)
        with pytest.raises(ValueError):
            cloud.centroid_formula = "kahan"
    rng = np.random.default_rng(41)
    pts = np.round(rng.uniform(0, 12, (30000, 3)), 3) + [476850.0, 5429100.0, 312.0]
    pts[5000:9000] = pts[1000:5000]  # duplicated points: exact ties
    x, y, z = (np.ascontiguousarray(pts[:, k]) for k in range(3))
    cloud = E.VoxelCloud.build(x, y, z, 0.5, [0, 0, 0])
    g = O.group_by_voxel(*O.voxelize(x, y, z, [0, 0, 0], 0.5))
    np.testing.assert_array_equal(_np(cloud.centroid_reference_formula()).T, O.center_of_gravity_reference_bitwise(x, y, z, g))
    np.testing.assert_array_equal(_np(cloud.std_reference_formula()).T, O.std_reference_bitwise(x, y, z, g))


@pytest.mark.parametrize("name", ["utm_offset_vs0.25", "uniform5000_vs0.25", "plane500_vs1.0", "vapc_in_every50th_vs0.25", "shifted_origin_vs0.3"])
def test_cov_reference_formula_bit_for_bit(E, name):
    """vapc_cov_reference_formula == the cov_** columns of the unmodified reference, bit for bit (UTM-sized coordinates included,
    where those columns are mostly rounding noise) — and == the oracle's Kahan restatement on a larger seeded cloud."""
    gold = P.load_golden(name)
    x, y, z = (np.asarray(gold[k], np.float64) for k in ("in_X", "in_Y", "in_Z"))
    vs, origin = float(gold["in_voxel_size"]), np.asarray(gold["in_origin"], float).tolist()
    for bucketed in (True, False):
        cloud = E.VoxelCloud.build(x, y, z, vs, origin, bucketed=bucketed)
        got = _np(cloud.covariance_reference_formula())[:, _u32(cloud.point2voxel)]
        for k, c in enumerate(("cov_xx", "cov_xy", "cov_xz", "cov_yy", "cov_yz", "cov_zz")):
            np.testing.assert_array_equal(got[k], np.real(np.asarray(gold["pt_" + c])), err_msg=f"{name} {c} bucketed={bucketed}")
    if name == "utm_offset_vs0.25":
        xs, ys, zs = make_cloud("utm", 30000, 77)
        cloud = E.VoxelCloud.build(xs, ys, zs, 0.5, [0, 0, 0])
        g = O.group_by_voxel(*O.voxelize(xs, ys, zs, [0, 0, 0], 0.5))
        np.testing.assert_array_equal(_np(cloud.covariance_reference_formula()).T, O.covariance_reference_bitwise(xs, ys, zs, g))


def test_reduce_lanes_per_voxel(E):
    """vapc_reduce_moments sums a voxel with 1, 2 or 4 lanes (chosen from n / V; forced here): integers identical, moments equal to
    rounding, and every variant bit-reproducible."""
    from vapc_b200 import _lib as L

    try:
        for kind, n, vs, seed in (("uniform", 400_000, 0.5, 21), ("clustered", 300_000, 0.25, 22), ("duplicates", 50_000, 1.0, 23)):
            x, y, z = make_cloud(kind, n, seed)
            L.call("vapc_reduce_debug_lanes", 1)
            ref = E.VoxelCloud.build(x, y, z, vs, [0, 0, 0])
            for lanes in (2, 4):
                L.call("vapc_reduce_debug_lanes", lanes)
                a = E.VoxelCloud.build(x, y, z, vs, [0, 0, 0])
                b = E.VoxelCloud.build(x, y, z, vs, [0, 0, 0])
                for name in ("count", "first_idx", "voxel_keys", "start", "point2voxel"):
                    assert torch.equal(getattr(a, name), getattr(ref, name)), (kind, lanes, name)
                assert torch.equal(a.mean3, b.mean3) and torch.equal(a.m2, b.m2), (kind, lanes, "reproducible")
                assert torch.allclose(a.mean3, ref.mean3, rtol=0, atol=1e-13 * vs), (kind, lanes)
                scale = ref.m2.abs().amax(0, keepdim=True).clamp(min=1e-300)
                assert bool(((a.m2 - ref.m2).abs() <= 1e-12 * scale + 1e-24).all()), (kind, lanes)
    finally:
        L.call("vapc_reduce_debug_lanes", 0)


def test_global_keys_and_histogram(E):
    """vapc_global_keys: the sorted local voxel keys re-expressed in a larger (global) key layout — same voxels, same order — and the
    histogram of their leading bits (few bins: runs of thousands of rows per bin; many bins: several runs per warp round)."""
    import ctypes as C

    from vapc_b200 import _lib as L

    for kind, n, vs, seed in (("uniform", 300_000, 0.1, 3), ("clustered", 200_000, 0.25, 4), ("uniform", 5_000, 2.0, 5)):
        x, y, z = make_cloud(kind, n, seed)
        cloud = E.VoxelCloud.build(x, y, z, vs, [0, 0, 0])
        lo = [float(np.min(c)) - 37.0 for c in (x, y, z)]
        hi = [float(np.max(c)) + 91.0 for c in (x, y, z)]
        g = E.grid_from_bounds(vs, [0, 0, 0], lo + hi)
        vx, vy, vz = (t.to(torch.int64) for t in cloud.voxel_coords())
        expect = ((vx - g.vmin[0]) << (g.bits[1] + g.bits[2])) | ((vy - g.vmin[1]) << g.bits[2]) | (vz - g.vmin[2])
        assert bool((expect[1:] > expect[:-1]).all()), "the global layout keeps the order"
        for hb in (4, 12, 16):
            hb = min(hb, g.total_bits)
            keys = torch.empty(cloud.nvoxels, dtype=torch.int64, device="cuda")
            hist = torch.empty(1 << hb, dtype=torch.int32, device="cuda")
            L.call("vapc_global_keys", E._ptr(cloud.voxel_keys), cloud.nvoxels, C.byref(cloud.grid), C.byref(g), hb, E._ptr(keys), E._ptr(hist), E._stream())
            assert torch.equal(keys, expect), (kind, hb)
            ref = torch.bincount(expect >> (g.total_bits - hb), minlength=1 << hb)
            assert torch.equal(hist.to(torch.int64), ref), (kind, hb)


def test_reproducible_bitwise(E):
    """No atomics in the reductions: two runs give bit-identical tables — and so does the path that keeps the
    records in input order and sorts all key bits (bucketed=False)."""
    x, y, z = make_cloud("clustered", 150000, 13)
    a = E.VoxelCloud.build(x, y, z, 0.25, [0, 0, 0])
    b = E.VoxelCloud.build(x, y, z, 0.25, [0, 0, 0])
    c = E.VoxelCloud.build(x, y, z, 0.25, [0, 0, 0], bucketed=False)
    for name in ("mean3", "m2", "count", "first_idx", "point2voxel", "idx_sorted", "keys_sorted", "voxel_keys", "start"):
        assert torch.equal(getattr(a, name), getattr(b, name)), name
        assert torch.equal(getattr(a, name), getattr(c, name)), name + " (bucketed vs input-order records)"


def _cell_sort_cloud(kind, seed):
    rng = np.random.default_rng(seed)
    if kind == "low9":  # 8 + 5 + 4 = 17 key bits: 9 below the bucket digit, the smallest cell tables
        ext, n = (25.6, 3.2, 1.6), 150_000
    elif kind == "low13":
        ext, n = (25.6, 25.6, 3.2), 400_000
    elif kind == "low16":  # 9 + 9 + 6 = 24 key bits like the benchmark grid: 2^16 cells per bucket
        ext, n = (50.0, 50.0, 4.0), 1_500_000
    else:
        raise ValueError(kind)
    pts = np.stack([rng.integers(0, int(e / 1e-4), n) * 1e-4 for e in ext], axis=1)
    # hot cells of 17 .. ~120 points (the 32-wide network and the warp-ranked cells), scattered through the input
    hot = []
    for cnt in (17, 20, 31, 32, 33, 40, 64, 96, 100, 110):
        for _ in range(6):
            c = np.floor(rng.uniform(0, 1, 3) * np.array(ext) / 0.1) * 0.1 + 0.05
            hot.append(c + rng.uniform(-0.04, 0.04, (cnt, 3)))
    pts = np.vstack([pts] + hot)
    pts = np.round(pts[rng.permutation(len(pts))], 4)
    return tuple(np.ascontiguousarray(pts[:, k]) for k in range(3))


@pytest.mark.parametrize("kind", ["low9", "low13", "low16"])
def test_cell_sort_equals_radix(E, kind):
    """vapc_sort_cells_segmented (dense grids) gives what the stable radix passes give, bit for bit: sorted keys, positions
    (input order inside a voxel), voxel count, and with them every table column."""
    x, y, z = _cell_sort_cloud(kind, 7)
    a = E.VoxelCloud.build(x, y, z, 0.1, [0, 0, 0], cell_sort=True)
    b = E.VoxelCloud.build(x, y, z, 0.1, [0, 0, 0], cell_sort=False)
    assert a.sorted_by == "cells" and b.sorted_by == "radix"
    assert a.nvoxels == b.nvoxels
    assert int(_u32(a.count).max()) >= 100
    for name in ("keys_sorted", "pos_sorted", "idx_sorted", "voxel_keys", "start", "count", "first_idx", "mean3", "m2", "point2voxel"):
        assert torch.equal(getattr(a, name), getattr(b, name)), name
    c = E.VoxelCloud.build(x, y, z, 0.1, [0, 0, 0], cell_sort=True)  # and run to run (the atomics only hand out slots)
    assert torch.equal(a.pos_sorted, c.pos_sorted) and torch.equal(a.m2, c.m2)


def test_cell_sort_falls_back_on_crowded_cells(E):
    """A cell with more than 128 points: the per-cell ordering is not the cell sort's job, status bit 2 -> radix passes."""
    x, y, z = _cell_sort_cloud("low13", 8)
    x[:500], y[:500], z[:500] = 3.33, 4.44, 1.11
    a = E.VoxelCloud.build(x, y, z, 0.1, [0, 0, 0], cell_sort=True)
    b = E.VoxelCloud.build(x, y, z, 0.1, [0, 0, 0], cell_sort=False)
    assert a.sorted_by == "radix"
    for name in ("keys_sorted", "pos_sorted", "voxel_keys", "count", "m2", "point2voxel"):
        assert torch.equal(getattr(a, name), getattr(b, name)), name


def test_bucketed_records(E):
    """vapc_pack_records_bucketed: buckets are the leading key bits in order, stable inside a bucket, every record
    carries its point and its original index."""
    import ctypes as C
    from vapc_b200 import _lib as L

    for kind, n, vs in (("uniform", 100_003, 0.05), ("clustered", 50_000, 0.5), ("utm", 30_000, 0.05), ("uniform", 7, 10.0), ("wide", 60_000, 0.1),
                        ("huge", 20_000, 0.001)):
        if kind in ("wide", "huge"):  # 2 km x 2 km x 50 m: 39 key bits at 0.1 m (32-bit low keys inside the buckets), 59 at 1 mm
            rng = np.random.default_rng(n)
            x, y, z = (rng.integers(0, int(e / 1e-4), n) * 1e-4 for e in (2000.0, 2000.0, 50.0))
        else:
            x, y, z = make_cloud(kind, n, seed=n)
        dx, dy, dz = (torch.from_numpy(a).cuda() for a in (x, y, z))
        g = E.grid_from_bounds(vs, [0, 0, 0], E.coordinate_bounds(dx, dy, dz))
        kd = torch.int32 if g.key_bytes == 4 else torch.int64
        keys_b, keys_o = torch.empty(n, dtype=kd, device="cuda"), torch.empty(n, dtype=kd, device="cuda")
        rec = torch.empty((n, 4), dtype=torch.float64, device="cuda")
        bstart = torch.empty(257, dtype=torch.int32, device="cuda")
        status = torch.zeros(1, dtype=torch.int32, device="cuda")
        wsb = L.raw("vapc_bucket_workspace_bytes")(n)
        ws = torch.empty(wsb, dtype=torch.uint8, device="cuda")
        bits, bkb = C.c_int32(0), C.c_int32(0)
        L.call("vapc_pack_records_bucketed", E._ptr(dx), E._ptr(dy), E._ptr(dz), n, C.byref(g), E._ptr(keys_b), E._ptr(rec), E._ptr(keys_o),
               E._ptr(bstart), C.byref(bits), C.byref(bkb), E._ptr(status), E._ptr(ws), wsb, E._stream())
        torch.cuda.synchronize()
        assert int(status.item()) == 0
        mb = bits.value
        assert mb == min(8, g.bits[0])
        ut = np.uint32 if g.key_bytes == 4 else np.uint64
        ko = _np(keys_o).view(ut).astype(np.uint64)
        vx, vy, vz = O.voxelize(x, y, z, [0, 0, 0], vs)
        ref = ((vx - g.vmin[0]).astype(np.uint64) << np.uint64(g.bits[1] + g.bits[2])) | ((vy - g.vmin[1]).astype(np.uint64) << np.uint64(g.bits[2])) \
            | (vz - g.vmin[2]).astype(np.uint64)
        np.testing.assert_array_equal(ko, ref)
        bucket = (ko >> np.uint64(g.total_bits - mb)).astype(np.int64)
        order = np.argsort(bucket, kind="stable")
        low_bits = g.total_bits - mb
        if kind == "wide":
            assert g.key_bytes == 8 and bkb.value == 4, "39-bit grid: the bucketed keys are the 31 low bits as 32-bit words"
        if kind == "huge":
            assert g.key_bytes == 8 and bkb.value == 8
        if bkb.value == 4 and g.key_bytes == 8:
            got = _np(keys_b).view(np.uint32)[:n].astype(np.uint64)
            np.testing.assert_array_equal(got, ko[order] & np.uint64((1 << low_bits) - 1))
            full = torch.empty(n, dtype=torch.int64, device="cuda")  # low parts are still in bucket order: expansion restores the keys
            L.call("vapc_expand_bucket_keys", E._ptr(keys_b), n, E._ptr(bstart), mb, low_bits, E._ptr(full), E._stream())
            np.testing.assert_array_equal(_np(full).view(np.uint64), ko[order])
        else:
            assert bkb.value == g.key_bytes
            np.testing.assert_array_equal(_np(keys_b).view(ut).astype(np.uint64), ko[order])
        r = _np(rec)
        np.testing.assert_array_equal(r[:, 3].copy().view(np.int64) & 0xFFFFFFFF, order)  # (32-bit keys ride in the upper half)
        np.testing.assert_array_equal(r[:, 0], x[order])
        np.testing.assert_array_equal(r[:, 1], y[order])
        np.testing.assert_array_equal(r[:, 2], z[order])
        nb = 1 << mb
        np.testing.assert_array_equal(_u32(bstart)[: nb + 1], np.concatenate([[0], np.cumsum(np.bincount(bucket, minlength=nb))]))


@pytest.mark.parametrize("n", [3_000_000])
def test_properties_at_scale(E, n):
    """Size-independent properties on a cloud too big for the per-voxel Python oracle."""
    rng = np.random.default_rng(21)
    x = rng.integers(0, 500000, n) * 1e-4
    y = rng.integers(0, 500000, n) * 1e-4
    z = rng.integers(0, 40000, n) * 1e-4
    vs = 0.1
    cloud = E.VoxelCloud.build(x, y, z, vs, [0, 0, 0])
    ks = _np(cloud.keys_sorted).view(np.uint32 if cloud.grid.key_bytes == 4 else np.uint64)
    assert np.all(ks[1:] >= ks[:-1]), "sortedness"
    idx = _u32(cloud.idx_sorted)
    assert np.array_equal(np.sort(idx), np.arange(n)), "permutation"
    count = _u32(cloud.count)
    assert count.sum() == n, "checksum of counts"
    vx, vy, vz = O.voxelize(x, y, z, [0, 0, 0], vs)
    packed = (vx * 1_000_000 + vy) * 1_000 + vz
    uniq, inv, cnt = np.unique(packed, return_inverse=True, return_counts=True)
    assert cloud.nvoxels == len(uniq)
    np.testing.assert_array_equal(count, cnt)
    np.testing.assert_array_equal(_u32(cloud.point2voxel), inv)
    st = cloud.stats(cog=True, cov=True, eig=True)
    cog = _np(st.cog).T
    for k, c in enumerate((x, y, z)):  # centroid * count == sum (linearity)
        s = np.bincount(inv, weights=c, minlength=len(uniq))
        np.testing.assert_allclose(cog[:, k] * cnt, s, rtol=1e-12)
    eig, cov = _np(st.eig), _np(st.cov)
    ok = cnt > 1
    np.testing.assert_allclose(eig[:, ok].sum(0), cov[0, ok] + cov[3, ok] + cov[5, ok], rtol=1e-9, atol=1e-15)  # trace
    assert np.all(eig[0, ok] >= eig[1, ok]) and np.all(eig[1, ok] >= eig[2, ok]) and np.all(eig[2, ok] >= 0)


def test_host_stream_pipeline_matches_direct(E):
    """The overlapped host-buffer pipeline returns, for every cloud, exactly the table of a direct call."""
    from vapc_b200.pipeline import HostStreamPipeline

    clouds = [make_cloud("uniform" if i % 2 else "clustered", 40000 + 7000 * i, seed=100 + i) for i in range(6)]
    pinned = [tuple(torch.from_numpy(c).pin_memory() for c in cl) for cl in clouds]
    pipe = HostStreamPipeline(0.3, [0.0, 0.0, 0.0])
    results = []
    for res in pipe.run(pinned):
        results.append({k: v.clone() for k, v in res.items() if isinstance(v, torch.Tensor)})
    assert len(results) == len(clouds)
    for cl, res in zip(clouds, results):
        direct = E.VoxelCloud.build(*cl, 0.3, [0.0, 0.0, 0.0])
        st = direct.stats(density=True, cog=True, std=True, cov=True, eig=True, eig_n=True, feat=True)
        assert torch.equal(res["voxel_keys"], direct.voxel_keys.cpu())
        assert torch.equal(res["count"], direct.count.cpu())
        for name in ("density", "cog", "std", "cov", "eig", "eig_n", "feat"):
            a, b = res[name], getattr(st, name).cpu()
            assert torch.equal(torch.isnan(a), torch.isnan(b)) and torch.equal(torch.nan_to_num(a), torch.nan_to_num(b)), name
    assert pipe.h2d_bytes == sum(3 * 8 * len(c[0]) for c in clouds)


# --------------------------------------------------------------------------------------
@pytest.mark.parametrize("n,scales,offsets,vs", [(200_000, [1e-4, 1e-4, 1e-4], [0.0, 0.0, 0.0], 0.1),
                                                 (150_001, [0.00025, 0.00025, 0.001], [476850.0, 5429100.0, 312.0], 0.25),
                                                 (70_000, [0.01, 0.01, 0.01], [-500.0, 20.0, -3.0], 0.05)])
def test_las_integer_input_is_bit_identical(E, n, scales, offsets, vs):
    """The cloud as int32 LAS coordinates (12 B/point) gives the same bits as its float64 columns X * scale + offset
    (what laspy hands the reference, datahandler.py:72-95): coordinates, keys, every row and moment of the voxel table."""
    rng = np.random.default_rng(n)
    q = [rng.integers(-2_000_000, 2_000_000, n).astype(np.int32), rng.integers(0, 1_500_000, n).astype(np.int32),
         rng.integers(-30_000, 90_000, n).astype(np.int32)]
    cols = [a.astype(np.float64) * s + o for a, s, o in zip(q, scales, offsets)]  # numpy: two rounded operations
    dev = [_np(t) for t in E.las_coordinates(*q, scales, offsets)]
    for a, b in zip(dev, cols):
        np.testing.assert_array_equal(a.view(np.int64), b.view(np.int64))
    origin = [offsets[0], offsets[1] - 1.0, 0.0]
    ref = E.VoxelCloud.build(*cols, vs, origin)
    got = E.VoxelCloud.build(None, None, None, vs, origin, las=(*q, scales, offsets))
    assert got.nvoxels == ref.nvoxels and got.grid.total_bits == ref.grid.total_bits and list(got.grid.vmin) == list(ref.grid.vmin)
    for name in ("voxel_keys", "start", "count", "first_idx", "mean3", "m2", "point2voxel", "idx_sorted", "x", "y", "z"):
        a, b = getattr(got, name), getattr(ref, name)
        assert torch.equal(a.view(torch.int64) if a.dtype == torch.float64 else a, b.view(torch.int64) if b.dtype == torch.float64 else b), name
    lean = E.VoxelCloud.build(None, None, None, vs, origin, las=(*q, scales, offsets), keep_columns=False, want_point2voxel=False)
    assert torch.equal(lean.m2.view(torch.int64), ref.m2.view(torch.int64)) and lean.x is None
    with pytest.raises(TypeError):
        E.VoxelCloud.build(None, None, None, vs, origin, las=(q[0].astype(np.int64), q[1], q[2], scales, offsets))


def test_run_lengths_and_mode_from_runs(E):
    """vapc_run_lengths (weighted run-length encoding of sorted keys) and vapc_mode_count_runs against numpy, and the runs
    route == the per-point route of mode / mode_count."""
    import ctypes as C

    from vapc_b200 import _lib as L

    rng = np.random.default_rng(21)
    for n, hi in ((1, 5), (5000, 40), (300_000, 20_000), (1_000_003, 7)):
        keys = np.sort(rng.integers(0, hi, n)).astype(np.int64)
        w = rng.integers(1, 1000, n).astype(np.int32)
        for weights in (None, w):
            rk, rw = E.run_lengths(torch.from_numpy(keys).cuda(), None if weights is None else torch.from_numpy(weights).cuda())
            uk, inv = np.unique(keys, return_inverse=True)
            np.testing.assert_array_equal(_np(rk), uk)
            np.testing.assert_array_equal(_np(rw), np.bincount(inv, weights=None if weights is None else weights.astype(np.float64)).astype(np.int64))
    # mode / mode_count from runs == from points
    x, y, z = make_cloud("clustered", 120000, 3)
    cloud = E.VoxelCloud.build(x, y, z, 0.5, [0, 0, 0])
    attr = rng.integers(0, 12, len(x)).astype(np.float64)
    mode, counts = cloud.mode_count(attr, [0.2], want_mode=True)
    attr_bits = 4
    comp = (_u32(cloud.point2voxel) << attr_bits) | attr.astype(np.int64)
    ks, _ = E.sort_pairs(torch.from_numpy(comp).cuda(), int(cloud.nvoxels - 1).bit_length() + attr_bits)
    rk, rw = E.run_lengths(ks)
    m2 = torch.empty(cloud.nvoxels, dtype=torch.int64, device="cuda")
    c2 = torch.empty(cloud.nvoxels, dtype=torch.int64, device="cuda")
    L.call("vapc_mode_count_runs", C.c_void_p(rk.data_ptr()), C.c_void_p(rw.data_ptr()), rk.numel(), attr_bits, cloud.nvoxels, 0.2,
           C.c_void_p(m2.data_ptr()), C.c_void_p(c2.data_ptr()), C.c_void_p(torch.cuda.current_stream().cuda_stream))
    assert torch.equal(m2, mode) and torch.equal(c2, counts[0.2])
