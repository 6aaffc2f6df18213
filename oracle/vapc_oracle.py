"""CPU oracle for the voxelise + per-voxel statistics path of 3dgeo-heidelberg/vapc.

THIS IS TEST INFRASTRUCTURE, NOT PRODUCT CODE.  Only `tests/`, `__graft_entry__.smoke()`
and `bench.py`'s `cpu_baseline` / `--impl reference` legs may import it; the product
package `vapc_b200` never does (it fails loudly when the CUDA library is missing).

It is a plain numpy restatement of what the reference computes, organised around a
*voxel table* (one row per occupied voxel, rows in lexicographic (vx, vy, vz) order) plus a
point->voxel map, instead of the reference's one-row-per-point pandas DataFrame.  Every
function cites the reference lines it restates (paths relative to /root/reference).

Pinning: `tests/test_oracle_golden.py` checks every function here against the fixtures in
`tests/golden/*.npz`, which `oracle/make_golden.py` produced by running the UNMODIFIED
reference in the build container (numpy 2.3.5 / pandas 3.0.2 / scipy 1.18.1); the
reference's own unit-test known answers (tests/test_voxel.py, tests/test_compute.py) are
replayed against it as well.  Parity is therefore pinned, except where stated below.

Summation: the reference's group sums run inside pandas (Kahan-compensated fp64,
`pandas._libs.groupby.group_sum/group_mean`).  The oracle accumulates in x87 long double
and rounds once, which agrees with a compensated sum to the last bit in all but
pathological cases; fixtures confirm 0-1 ulp agreement.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field

import numpy as np

LD = np.longdouble

FEATURE_NAMES = (
    "Sum_of_Eigenvalues",
    "Omnivariance",
    "Eigenentropy",
    "Anisotropy",
    "Planarity",
    "Linearity",
    "Surface_Variation",
    "Sphericity",
)
COV_NAMES = ("cov_xx", "cov_xy", "cov_xz", "cov_yx", "cov_yy", "cov_yz", "cov_zx", "cov_zy", "cov_zz")


# --------------------------------------------------------------------------------------
# a2  Vapc.voxelize                                             src/vapc/vapc.py:188-207
# --------------------------------------------------------------------------------------
def voxelize(x, y, z, origin, voxel_size):
    """voxel_d = floor((coord_d - origin_d) / voxel_size).astype(int64)  (vapc.py:199-201).

    fp64 subtract, true fp64 divide, floor, cast.  No reciprocal multiply, no FMA.
    """
    vs = float(voxel_size)
    out = []
    for c, o in zip((x, y, z), origin):
        c = np.asarray(c, dtype=np.float64)
        out.append(np.floor((c - o) / vs).astype(np.int64))
    return out[0], out[1], out[2]


@dataclass
class Grouping:
    """Result of grouping points by their voxel triple (the pandas groupby of vapc.py:724,
    :843, :877, :959, :1081 restated as a stable lexicographic sort)."""

    order: np.ndarray  # permutation: sorted position -> original point index (stable)
    starts: np.ndarray  # [V] first sorted position of each voxel
    counts: np.ndarray  # [V] int64 points per voxel
    point2voxel: np.ndarray  # [N] voxel row of every point
    vx: np.ndarray  # [V] int64
    vy: np.ndarray
    vz: np.ndarray
    first_index: np.ndarray = field(default=None)  # [V] lowest original index in the voxel

    @property
    def nvoxels(self):
        return int(self.starts.shape[0])


def group_by_voxel(vx, vy, vz) -> Grouping:
    """Group points by (vx, vy, vz); voxel rows come out in lexicographic (vx, vy, vz)
    order — the order the reference emits for `return_at="center_of_voxel"`
    (tests/test_voxel.py:217-229) — and points inside a voxel keep original order."""
    n = vx.shape[0]
    if n == 0:
        e = np.zeros(0, np.int64)
        return Grouping(e, e, e, e, e, e, e, e)
    order = np.lexsort((vz, vy, vx))  # stable; last key is primary
    sx, sy, sz = vx[order], vy[order], vz[order]
    head = np.ones(n, dtype=bool)
    head[1:] = (sx[1:] != sx[:-1]) | (sy[1:] != sy[:-1]) | (sz[1:] != sz[:-1])
    starts = np.flatnonzero(head).astype(np.int64)
    counts = np.diff(np.append(starts, n)).astype(np.int64)
    vid_sorted = np.cumsum(head) - 1
    p2v = np.empty(n, np.int64)
    p2v[order] = vid_sorted
    return Grouping(order.astype(np.int64), starts, counts, p2v, sx[starts], sy[starts], sz[starts], order[starts].astype(np.int64))


def _group_sum(values, g: Grouping):
    """Per-voxel sum, accumulated in long double and rounded once (see module docstring)."""
    if g.nvoxels == 0:
        return np.zeros(0, np.float64)
    return np.add.reduceat(np.asarray(values, dtype=np.float64)[g.order].astype(LD), g.starts).astype(np.float64)


def _group_sum_ld(values_ld_sorted, g: Grouping):
    return np.add.reduceat(values_ld_sorted, g.starts)


# --------------------------------------------------------------------------------------
# a4-a6  point_count / point_density / percentage_occupied      vapc.py:715-773
# --------------------------------------------------------------------------------------
def point_count(g: Grouping):
    """groupby(...)["X"].transform("size")  (vapc.py:724-725) -> int64 per voxel."""
    return g.counts.copy()


def point_density(g: Grouping, voxel_size):
    """point_count / voxel_size**3  (vapc.py:740)."""
    return g.counts / (voxel_size ** 3)


def percentage_occupied(g: Grouping):
    """round(V / bbox_cells * 100, 2) over the voxel-index bounding box (vapc.py:761-773)."""
    ex = int(g.vx.max()) - int(g.vx.min()) + 1
    ey = int(g.vy.max()) - int(g.vy.min()) + 1
    ez = int(g.vz.max()) - int(g.vz.min()) + 1
    return round(g.nvoxels / (ex * ey * ez) * 100, 2)


# --------------------------------------------------------------------------------------
# a7-a8  center_of_gravity / std_of_cog                         vapc.py:829-887
# --------------------------------------------------------------------------------------
def center_of_gravity(x, y, z, g: Grouping):
    """Per-voxel mean of X, Y, Z (groupby.transform("mean"), vapc.py:843-844)."""
    n = g.counts.astype(np.float64)
    return np.stack([_group_sum(c, g) / n for c in (x, y, z)], axis=1)


def std_of_cog(x, y, z, g: Grouping):
    """Per-voxel sample standard deviation (ddof=1) of X, Y, Z; NaN when n == 1
    (groupby.transform("std"), vapc.py:877-878).  Two-pass in long double."""
    out = np.full((g.nvoxels, 3), np.nan)
    nl = g.counts.astype(LD)
    for k, c in enumerate((x, y, z)):
        cs = np.asarray(c, np.float64)[g.order].astype(LD)
        mean = _group_sum_ld(cs, g) / nl
        dev = cs - np.repeat(mean, g.counts)
        ss = _group_sum_ld(dev * dev, g)
        with np.errstate(invalid="ignore", divide="ignore"):
            var = np.where(g.counts > 1, ss / (nl - 1), np.nan)
        out[:, k] = np.sqrt(var).astype(np.float64)
    return out


# --------------------------------------------------------------------------------------
# a12  covariance_matrix                                        vapc.py:929-1040
# --------------------------------------------------------------------------------------
_PAIRS = ((0, 0), (0, 1), (0, 2), (1, 1), (1, 2), (2, 2))  # xx xy xz yy yz zz


def covariance_reference_formula(x, y, z, g: Grouping):
    """The reference's raw-moment covariance: per point products in fp64 (vapc.py:950-955),
    per voxel sums (vapc.py:962-971), cov_ab = (S_ab - S_a*S_b/n)/(n-1), NaN when n <= 1
    (vapc.py:977-1006).  Returns [V, 6] in the order xx, xy, xz, yy, yz, zz.

    This formula is ill-conditioned when |coord| >> voxel_size; it is kept
    here verbatim because it is what the reference outputs."""
    c = [np.asarray(a, np.float64) for a in (x, y, z)]
    n = g.counts.astype(np.float64)
    s = [_group_sum(a, g) for a in c]
    out = np.full((g.nvoxels, 6), np.nan)
    with np.errstate(invalid="ignore", divide="ignore"):
        for k, (a, b) in enumerate(_PAIRS):
            sab = _group_sum(c[a] * c[b], g)
            if a == b:
                val = (sab - (s[a] ** 2) / n) / (n - 1)
            else:
                val = (sab - (s[a] * s[b]) / n) / (n - 1)
            out[:, k] = np.where(n > 1, val, np.nan)
    return out


def covariance_reference_bitwise(x, y, z, g: Grouping):
    """The reference's cov_* columns bit for bit (vapc.py:947-1006): per-point products rounded to fp64 (:950-955), per-voxel
    sums as pandas' groupby sum forms them (`grouped[...].transform("sum")`, :962-971: Kahan-compensated, rows in frame order),
    then (S_ab - S_a S_b / n) / (n - 1), one rounding per operation (:977-1006).  [V, 6] in the order xx, xy, xz, yy, yz, zz.
    Pure-Python loop over the points: small clouds only.  tests/test_oracle_golden.py pins it to the golden columns with
    assert_array_equal."""
    c = [np.asarray(a, np.float64) for a in (x, y, z)]
    cols = c + [c[0] * c[0], c[1] * c[1], c[2] * c[2], c[0] * c[1], c[0] * c[2], c[1] * c[2]]
    lab = np.asarray(g.point2voxel, np.int64)
    s = np.zeros((9, g.nvoxels))
    comp = np.zeros((9, g.nvoxels))
    for j, col in enumerate(cols):
        sj, cj = s[j], comp[j]
        for v, l in zip(col.tolist(), lab.tolist()):
            y_ = v - cj[l]
            t = sj[l] + y_
            cj[l] = (t - sj[l]) - y_
            if cj[l] != cj[l]:
                cj[l] = 0.0
            sj[l] = t
    n = g.counts.astype(np.float64)
    out = np.full((g.nvoxels, 6), np.nan)
    with np.errstate(invalid="ignore", divide="ignore"):
        for k, (a, b, ab) in enumerate(((0, 0, 3), (0, 1, 6), (0, 2, 7), (1, 1, 4), (1, 2, 8), (2, 2, 5))):
            out[:, k] = np.where(n > 1, (s[ab] - (s[a] * s[b]) / n) / (n - 1), np.nan)
    return out


def center_of_gravity_reference_bitwise(x, y, z, g: Grouping):
    """The reference's cog_* columns bit for bit (vapc.py:843-850): pandas' groupby mean = the Kahan-compensated sum of the voxel's
    values in frame order (the same recurrence as its groupby sum) divided by the count.  [V, 3].  Pure-Python loop over the points:
    small clouds only.  tests/test_oracle_golden.py pins it — and the distance / min_distance columns that follow from it with numpy's
    own operation order (vapc.py:790-825) — to the golden columns with assert_array_equal."""
    lab = np.asarray(g.point2voxel, np.int64)
    out = np.empty((g.nvoxels, 3))
    for j, col in enumerate((x, y, z)):
        sj, cj = np.zeros(g.nvoxels), np.zeros(g.nvoxels)
        for v, l in zip(np.asarray(col, np.float64).tolist(), lab.tolist()):
            y_ = v - cj[l]
            t = sj[l] + y_
            cj[l] = (t - sj[l]) - y_
            if cj[l] != cj[l]:
                cj[l] = 0.0
            sj[l] = t
        out[:, j] = sj / g.counts.astype(np.float64)
    return out


def std_reference_bitwise(x, y, z, g: Grouping):
    """The reference's std_* columns bit for bit (vapc.py:873-880): pandas' groupby std = Welford's recurrence over the voxel's values
    in frame order, M2 / (n - 1), square root.  [V, 3], NaN for single-point voxels.  Pure-Python loop: small clouds only; pinned to
    the golden columns with assert_array_equal (tests/test_oracle_golden.py)."""
    lab = np.asarray(g.point2voxel, np.int64)
    out = np.empty((g.nvoxels, 3))
    for j, col in enumerate((x, y, z)):
        n, mean, m2 = np.zeros(g.nvoxels), np.zeros(g.nvoxels), np.zeros(g.nvoxels)
        for v, l in zip(np.asarray(col, np.float64).tolist(), lab.tolist()):
            n[l] += 1.0
            old = mean[l]
            mean[l] = old + (v - old) / n[l]
            m2[l] += (v - mean[l]) * (v - old)
        with np.errstate(invalid="ignore", divide="ignore"):
            out[:, j] = np.sqrt(np.where(n > 1, m2 / (n - 1), np.nan))
    return out


def covariance_exact(x, y, z, g: Grouping):
    """Two-pass centred covariance in long double (what np.cov computes, vapc.py:1055-1056,
    evaluated in extended precision).  Independent 'truth' for the H1 parity rule."""
    nl = g.counts.astype(LD)
    dev = []
    for c in (x, y, z):
        cs = np.asarray(c, np.float64)[g.order].astype(LD)
        mean = _group_sum_ld(cs, g) / nl
        dev.append(cs - np.repeat(mean, g.counts))
    out = np.full((g.nvoxels, 6), np.nan)
    with np.errstate(invalid="ignore", divide="ignore"):
        for k, (a, b) in enumerate(_PAIRS):
            ss = _group_sum_ld(dev[a] * dev[b], g)
            out[:, k] = np.where(g.counts > 1, ss / (nl - 1), np.nan).astype(np.float64)
    return out


def cov6_to_cov9(c6):
    """xx xy xz yy yz zz -> the 9 reference columns cov_xx..cov_zz (vapc.py:1008-1034)."""
    return c6[:, [0, 1, 2, 1, 3, 4, 2, 4, 5]]


# --------------------------------------------------------------------------------------
# a13  eigenvalues                                              vapc.py:1044-1090
# --------------------------------------------------------------------------------------
def eigenvalues(x, y, z, g: Grouping, faithful=True):
    """Per voxel: np.cov of the raw points (ddof=1) -> eigenvalues, sorted descending;
    NaN x3 when the covariance holds a NaN, i.e. n == 1 (vapc.py:1055-1065).

    faithful=True follows the reference literally (np.cov + np.linalg.eig per voxel, a
    Python loop); faithful=False is a batched np.linalg.eigvalsh on the same centred
    covariances for large parity cases (differences are at rounding level)."""
    pts = np.stack([np.asarray(a, np.float64) for a in (x, y, z)], axis=1)[g.order]
    out = np.full((g.nvoxels, 3), np.nan)
    if faithful:
        for v in range(g.nvoxels):
            n = int(g.counts[v])
            if n < 2:
                continue
            s = int(g.starts[v])
            cov = np.cov(pts[s : s + n].T)
            w = np.linalg.eig(cov)[0]
            w = np.real(w)
            out[v] = w[w.argsort()[::-1]]
        return out
    n = g.counts.astype(np.float64)
    mean = np.add.reduceat(pts, g.starts, axis=0) / n[:, None]
    dev = pts - np.repeat(mean, g.counts, axis=0)
    prods = np.einsum("ni,nj->nij", dev, dev)
    with np.errstate(invalid="ignore", divide="ignore"):
        cov = np.add.reduceat(prods, g.starts, axis=0) / (n - 1)[:, None, None]
    ok = g.counts > 1
    out[ok] = np.linalg.eigvalsh(cov[ok])[:, ::-1]
    return out


# --------------------------------------------------------------------------------------
# a14  geometric_features                                       vapc.py:1094-1146
# --------------------------------------------------------------------------------------
def geometric_features(eig):
    """The 8 eigenvalue features and the 3 normalised eigenvalues the reference leaves in
    the frame (vapc.py:1112-1145).  Returns (features [V, 8] in FEATURE_NAMES order,
    eig_n [V, 3])."""
    l1, l2, l3 = eig[:, 0], eig[:, 1], eig[:, 2]
    with np.errstate(invalid="ignore", divide="ignore"):
        total = l1 + l2 + l3
        en = np.stack([l1 / total, l2 / total, l3 / total], axis=1)

        def safe_log(v):  # vapc.py:1104-1106
            return np.where(v > 0, np.log(np.where(v > 0, v, 1.0)), 0)

        ssum = l1 + l2 + l3
        omni = (l1 * l2 * l3) ** (1 / 3)
        entropy = -(en[:, 0] * safe_log(en[:, 0]) + en[:, 1] * safe_log(en[:, 1]) + en[:, 2] * safe_log(en[:, 2]))
        aniso = (l1 - l3) / l1
        planar = (l2 - l3) / l1
        linear = (l1 - l2) / l1
        surf = l3 / ssum
        spher = l3 / l1
    return np.stack([ssum, omni, entropy, aniso, planar, linear, surf, spher], axis=1), en


# --------------------------------------------------------------------------------------
# a11  center_of_voxel / corner_of_voxel                        vapc.py:891-925
# --------------------------------------------------------------------------------------
def center_of_voxel(g: Grouping, origin, voxel_size):
    """((v * vs) + vs / 2) + origin, in that order (vapc.py:902-906)."""
    v = np.stack([g.vx, g.vy, g.vz], axis=1)
    return ((v * voxel_size) + voxel_size / 2) + np.asarray(origin)


def corner_of_voxel(g: Grouping, origin, voxel_size):
    """(v * vs) + origin (vapc.py:922-924)."""
    v = np.stack([g.vx, g.vy, g.vz], axis=1)
    return (v * voxel_size) + np.asarray(origin)


# --------------------------------------------------------------------------------------
# a9-a10  distance / closest_to_center_of_gravity               vapc.py:776-825
# --------------------------------------------------------------------------------------
def distance_to_center_of_gravity(x, y, z, g: Grouping, cog):
    """Per POINT: sqrt((X-cog_x)**2 + (Y-cog_y)**2 + (Z-cog_z)**2) with the voxel's centroid
    broadcast to its points (vapc.py:790-794); fp64, squares are plain products, sum left
    to right."""
    c = cog[g.point2voxel]
    dx = np.asarray(x, np.float64) - c[:, 0]
    dy = np.asarray(y, np.float64) - c[:, 1]
    dz = np.asarray(z, np.float64) - c[:, 2]
    return np.sqrt(dx * dx + dy * dy + dz * dz)


def closest_to_center_of_gravity(dist, g: Grouping):
    """Per voxel: min distance (vapc.py:819-821) and the winning point.  The reference
    keeps every point with distance == min_distance (vapc.py:1195); we return the minimum,
    the lowest original index attaining it (the north_star tie-break), and the per-point
    `is_min` flag that reproduces the reference's multi-row behaviour."""
    ds = dist[g.order]
    dmin = np.minimum.reduceat(ds, g.starts) if g.nvoxels else np.zeros(0)
    is_min_sorted = ds == np.repeat(dmin, g.counts)
    pos = np.where(is_min_sorted, np.arange(ds.shape[0]), ds.shape[0])
    first = np.minimum.reduceat(pos, g.starts) if g.nvoxels else np.zeros(0, np.int64)
    winner = g.order[first]
    is_min = np.empty(ds.shape[0], bool)
    is_min[g.order] = is_min_sorted
    return dmin, winner.astype(np.int64), is_min


# --------------------------------------------------------------------------------------
# a16  per-attribute statistics                                 vapc.py:303-472
# --------------------------------------------------------------------------------------
def attribute_statistic(attr, g: Grouping, stat: str):
    """One statistic of one attribute per voxel, following the per-voxel Python loop of
    vapc.py:387-439 (numpy calls on each voxel's slice, in sorted-by-voxel order).

    mode  = argmax(bincount(int(x)))  -> the smallest most frequent value (vapc.py:407-410)
    mode_count,p = number of distinct values whose share count/n >= p (vapc.py:413-433)
    std/var use numpy's ddof=0 (vapc.py:395-398)."""
    a = np.asarray(attr)[g.order]
    out = []
    thr = None
    if "mode_count" in stat:
        try:
            _, p = stat.split(",")
            thr = float(p)
        except ValueError as exc:
            raise ValueError(f"Invalid 'mode_count' specification: '{stat}'") from exc
    for v in range(g.nvoxels):
        s = int(g.starts[v])
        sl = a[s : s + int(g.counts[v])]
        if stat == "mean":
            out.append(sl.mean())
        elif stat == "std":
            out.append(sl.std())
        elif stat == "var":
            out.append(sl.var())
        elif stat == "median":
            out.append(np.median(sl))
        elif stat == "min":
            out.append(sl.min())
        elif stat == "max":
            out.append(sl.max())
        elif stat == "sum":
            out.append(sl.sum())
        elif stat == "mode":
            out.append(int(np.argmax(np.bincount(sl.astype(int)))))
        elif thr is not None:
            counts = np.bincount(sl.astype(int))
            total = counts.sum()
            counts = counts[counts != 0]
            out.append(int(np.sum(counts / total >= thr)))
        else:
            raise ValueError(f"unknown statistic {stat}")
    return np.array(out)


# --------------------------------------------------------------------------------------
# a17-a18  voxel buffer / mask                                  vapc.py:495-609
# --------------------------------------------------------------------------------------
def voxel_buffer(vx, vy, vz, buffer_size=1):
    """Unique voxels of a (mask) cloud, each expanded by all (2b+1)^3 integer offsets,
    de-duplicated keeping first occurrence (vapc.py:507-525).  Returns [M, 3] int64 in the
    reference's row order (voxel-major, offsets in meshgrid 'ij' order)."""
    coords = np.stack([vx, vy, vz], axis=1)
    _, first = np.unique(coords, axis=0, return_index=True)
    coords = coords[np.sort(first)]  # drop_duplicates keeps first occurrences in order
    off = np.arange(-buffer_size, buffer_size + 1, dtype=np.int64)
    grid = np.stack(np.meshgrid(off, off, off, indexing="ij"), axis=-1).reshape(-1, 3)
    expanded = (coords[:, None, :] + grid[None, :, :]).reshape(-1, 3)
    _, first = np.unique(expanded, axis=0, return_index=True)
    return expanded[np.sort(first)]


def mask_membership(vx, vy, vz, mvx, mvy, mvz):
    """Boolean per point: is its voxel triple in the mask's voxel set
    (df.index.isin(mask_values), vapc.py:594-598)."""
    mask = set(zip(mvx.tolist(), mvy.tolist(), mvz.tolist()))
    return np.fromiter(((a, b, c) in mask for a, b, c in zip(vx.tolist(), vy.tolist(), vz.tolist())), bool, count=len(vx))


# --------------------------------------------------------------------------------------
# a19  filter_attributes                                        vapc.py:1214-1270
# --------------------------------------------------------------------------------------
_OPS = {
    "equal_to": np.equal, "==": np.equal,
    "not_equal_to": np.not_equal, "!=": np.not_equal,
    "greater_than": np.greater, ">": np.greater,
    "greater_than_or_equal_to": np.greater_equal, ">=": np.greater_equal,
    "less_than": np.less, "<": np.less,
    "less_than_or_equal_to": np.less_equal, "<=": np.less_equal,
}


def filter_mask(values, min_max_eq: str, filter_value):
    """Row-selection mask of filter_attributes (vapc.py:1256-1270); ValueError on an
    unknown operator, operator names are case-insensitive."""
    op = _OPS.get(min_max_eq.lower())
    if op is None:
        raise ValueError("Filter invalid")
    return op(np.asarray(values), filter_value)


# --------------------------------------------------------------------------------------
# a15  reduce_to_voxels                                         vapc.py:1150-1210
# --------------------------------------------------------------------------------------
def reduce_to_voxels(x, y, z, g: Grouping, return_at, origin, voxel_size, extra=None):
    """Rows the reference emits from reduce_to_voxels, as (XYZ [R, 3], source point index
    [R]) in the reference's output order (sorted by the output X, Y, Z floats,
    vapc.py:1209).  For centre/corner/centroid modes every voxel yields one row whose other
    columns are those of its first point (drop_duplicates keeps the first row,
    vapc.py:1208); for closest_to_center_of_gravity all exactly tied points are kept and
    identical XYZ rows are merged (vapc.py:1195, :1208)."""
    if return_at == "center_of_voxel":
        xyz = center_of_voxel(g, origin, voxel_size)
        src = g.first_index
    elif return_at == "corner_of_voxel":
        xyz = corner_of_voxel(g, origin, voxel_size)
        src = g.first_index
    elif return_at == "center_of_gravity":
        xyz = center_of_gravity(x, y, z, g)
        src = g.first_index
    elif return_at == "closest_to_center_of_gravity":
        cog = center_of_gravity(x, y, z, g)
        dist = distance_to_center_of_gravity(x, y, z, g, cog)
        _, _, is_min = closest_to_center_of_gravity(dist, g)
        src = np.flatnonzero(is_min)
        xyz = np.stack([np.asarray(a, np.float64)[src] for a in (x, y, z)], axis=1)
    else:
        raise ValueError(return_at)
    # drop_duplicates(subset=XYZ) keeps the first row per distinct XYZ, then the groupby
    # on (X, Y, Z) sorts rows by those floats.
    _, first = np.unique(xyz, axis=0, return_index=True)
    first = first[np.lexsort((xyz[first, 2], xyz[first, 1], xyz[first, 0]))]
    return xyz[first], np.asarray(src)[first]


# --------------------------------------------------------------------------------------
# a21-a25  bi-temporal change detection                         src/vapc/bi_vapc.py
# --------------------------------------------------------------------------------------
def chi2_sf3_one_minus_cdf(d):
    """p = 1 - chi2.cdf(d, df=3) (bi_vapc.py:118-119) in closed form:
    cdf = erf(sqrt(d/2)) - sqrt(2 d / pi) * exp(-d/2).  The reference forms 1 - cdf (not a
    survival function), so p is exactly 0 for d >~ 79, which drives change_type 2
    (bi_vapc.py:135); this mirrors that."""
    d = np.asarray(d, np.float64)
    out = np.full(d.shape, np.nan)
    ok = ~np.isnan(d)
    dd = np.where(ok, d, 0.0)
    erf = np.vectorize(math.erf, otypes=[np.float64])
    with np.errstate(over="ignore", invalid="ignore"):
        cdf = erf(np.sqrt(dd / 2)) - np.sqrt(2 * dd / math.pi) * np.exp(-dd / 2)
    cdf = np.where(dd <= 0, 0.0, cdf)
    out[ok] = (1 - cdf)[ok]
    return out


def join_epochs(v1, v2):
    """Inner join of two voxel tables on the voxel triple, left (epoch-1) order preserved
    (df1.add_suffix("_x").join(df2.add_suffix("_y"), how="inner"), bi_vapc.py:84), plus
    the two anti-joins (index.difference both ways, bi_vapc.py:30-31).
    v1, v2: [V, 3] int64 voxel coordinates.  Returns (i1, i2, only1, only2) row indices."""
    d2 = {t: i for i, t in enumerate(map(tuple, v2.tolist()))}
    d1 = {t: i for i, t in enumerate(map(tuple, v1.tolist()))}
    i1, i2 = [], []
    for i, t in enumerate(map(tuple, v1.tolist())):
        j = d2.get(t)
        if j is not None:
            i1.append(i)
            i2.append(j)
    only1 = sorted((t for t in d1 if t not in d2))  # Index.difference returns sorted
    only2 = sorted((t for t in d2 if t not in d1))
    return (np.array(i1, np.int64), np.array(i2, np.int64),
            np.array([d1[t] for t in only1], np.int64), np.array([d2[t] for t in only2], np.int64))


def mahalanobis_change(cog1, cov1_9, cog2, cov2_9, alpha=0.005):
    """compute_mahalanobis_distance (bi_vapc.py:92-138) on already-joined rows.

    cov*_9: [R, 9] row-major 3x3 (the sorted cov_* column order xx,xy,xz,yx,...,zz is
    row-major, bi_vapc.py:101-107).  eps*I is added (bi_vapc.py:108-110), matrices are
    inverted, d1 = diff' inv(cov2) diff, d2 = diff' inv(cov1) diff (bi_vapc.py:116-117).
    Returns dict(d1, d2, p_value, significance, change_type)."""
    r = cog1.shape[0]
    eps = 1e-10
    cx = cov1_9.reshape(r, 3, 3) + np.eye(3) * eps
    cy = cov2_9.reshape(r, 3, 3) + np.eye(3) * eps
    diff = cog1 - cog2
    d1 = np.full(r, np.nan)
    d2 = np.full(r, np.nan)
    okx = ~np.isnan(cx).any(axis=(1, 2))
    oky = ~np.isnan(cy).any(axis=(1, 2))
    if oky.any():
        d1[oky] = np.einsum("ij,ijk,ik->i", diff[oky], np.linalg.inv(cy[oky]), diff[oky])
    if okx.any():
        d2[okx] = np.einsum("ij,ijk,ik->i", diff[okx], np.linalg.inv(cx[okx]), diff[okx])
    p1 = chi2_sf3_one_minus_cdf(d1)
    p2 = chi2_sf3_one_minus_cdf(d2)
    with np.errstate(invalid="ignore"):
        sig = (p1 < alpha) | (p2 < alpha)
        p = np.where(p1 < alpha, p1, p2)
    change = np.ones(r, np.int64)
    change[~sig] = 0
    change[p == 0] = 2
    return dict(d1=d1, d2=d2, p_value=p, significance=sig.astype(np.int64), change_type=change)


def euclidean_distance(cog1, cog2):
    """sqrt(dx**2 + dy**2 + dz**2) of joined centroids (bi_vapc.py:198-207)."""
    d = cog1 - cog2
    return np.sqrt(d[:, 0] ** 2 + d[:, 1] ** 2 + d[:, 2] ** 2)


# --------------------------------------------------------------------------------------
# (f)4  Vapc.compute_clusters                                   src/vapc/vapc.py:1274-1357
# --------------------------------------------------------------------------------------
def compute_clusters(pts, cluster_distance, nr_of_neighbours=50, dtype=None):
    """The reference's region growing restated literally (vapc.py:1336-1357): one KD-tree query of the
    `nr_of_neighbours` nearest rows per row, in row order; rows within `cluster_distance` that carry no label yet
    open label `objectCounter`; otherwise every label met is rewritten to the lowest one (frame-wide) and the rows
    of the query take it.  pts: [n, k] rows of df[cluster_by].  Returns (cluster_id [n], cluster_size [n]); dtype
    follows the reference: that of the first cluster_by column (np.zeros_like, vapc.py:1338)."""
    from scipy.spatial import KDTree

    pts = np.asarray(pts)
    n = len(pts)
    oc = np.zeros(n, dtype=dtype if dtype is not None else pts.dtype)
    tree = KDTree(pts)
    counter = 1.0
    for point in pts:
        distances, indices = tree.query(point, nr_of_neighbours)
        rel = indices[np.where(distances <= cluster_distance)]
        if oc[rel].max() < 1:
            oc[rel] = counter
            counter += 1
        else:
            obj = np.unique(oc[rel])
            existing = obj[obj > 0]
            for e in existing:
                oc[oc == e] = existing.min()
            oc[rel] = existing.min()
    oids, cts = np.unique(oc, return_counts=True)
    size = cts[np.searchsorted(oids, oc)]
    return oc, size.astype(np.result_type(oc.dtype, np.int64))


def cluster_components(pts, cluster_distance, dtype=None):
    """What compute_clusters amounts to whenever no row has more than `nr_of_neighbours` rows (itself included)
    within `cluster_distance` — stated without the sequential sweep, as the GPU computes it:
      * the partition is the connected components of the graph "distance <= cluster_distance";
      * row i opens a new label iff no earlier row lies within two hops of it (the rows of query j and of query i
        intersect iff j is within two hops of i), labels are numbered 1, 2, ... in row order of those openings;
      * merging always keeps the lowest label, so a component ends with the label opened by its lowest row.
    Returns (cluster_id, cluster_size) like compute_clusters."""
    from scipy.sparse import coo_matrix
    from scipy.sparse.csgraph import connected_components
    from scipy.spatial import KDTree

    pts = np.asarray(pts)
    n = len(pts)
    tree = KDTree(pts.astype(np.float64))
    pairs = tree.query_pairs(cluster_distance * (1 + 1e-12), output_type="ndarray")
    if len(pairs):
        d = pts[pairs[:, 0]].astype(np.float64) - pts[pairs[:, 1]].astype(np.float64)
        pairs = pairs[np.sqrt((d * d).sum(axis=1)) <= cluster_distance]
    a = np.concatenate([pairs[:, 0], pairs[:, 1], np.arange(n)])
    b = np.concatenate([pairs[:, 1], pairs[:, 0], np.arange(n)])
    adj = coo_matrix((np.ones(len(a), np.int64), (a, b)), shape=(n, n)).tocsr()
    _, comp = connected_components(adj, directed=False)
    m1 = np.full(n, n, np.int64)
    np.minimum.at(m1, a, b)                      # lowest row in the closed neighbourhood
    m2 = np.full(n, n, np.int64)
    np.minimum.at(m2, a, m1[b])                  # lowest row within two hops
    opens = np.flatnonzero(m2 == np.arange(n))   # rows that open a label, ascending
    comp_first = np.full(comp.max() + 1, n, np.int64)
    np.minimum.at(comp_first, comp, np.arange(n))
    label = np.searchsorted(opens, comp_first[comp]) + 1
    oc = label.astype(dtype if dtype is not None else pts.dtype)
    size = np.bincount(comp)[comp]
    return oc, size.astype(np.result_type(oc.dtype, np.int64))


# --------------------------------------------------------------------------------------
# (f)4  mutual nearest neighbours of two centroid sets           src/vapc/bi_vapc.py:51-74, 173-186
# --------------------------------------------------------------------------------------
def mutual_nearest_neighbors(coords1, coords2):
    """bi_vapc.py:173-186: nearest row of set 2 for every row of set 1, kept where that row's nearest row of set 1
    is the row itself.  Returns (idx1, idx2, distance) of the mutual pairs in set-1 order; distance as
    merge_vapcs_with_closest_point recomputes it (bi_vapc.py:62-64)."""
    from scipy.spatial import KDTree

    coords1 = np.asarray(coords1, np.float64)
    coords2 = np.asarray(coords2, np.float64)
    _, indices = KDTree(coords2).query(coords1, k=1)
    _, back = KDTree(coords1).query(coords2[indices], k=1)
    mutual = np.arange(len(coords1)) == back
    idx1, idx2 = np.arange(len(coords1))[mutual], indices[mutual]
    dist, _ = KDTree(coords2[idx2]).query(coords1[idx1], k=1)
    return idx1, idx2, dist


# --------------------------------------------------------------------------------------
# convenience: the whole per-voxel table for one cloud
# --------------------------------------------------------------------------------------
def voxel_table(x, y, z, voxel_size, origin=(0.0, 0.0, 0.0), faithful_eig=True):
    """All per-voxel outputs of the hot path for one cloud, as a dict of arrays."""
    vx, vy, vz = voxelize(x, y, z, origin, voxel_size)
    g = group_by_voxel(vx, vy, vz)
    cog = center_of_gravity(x, y, z, g)
    eig = eigenvalues(x, y, z, g, faithful=faithful_eig)
    feat, eig_n = geometric_features(eig)
    return dict(
        grouping=g, voxel_x=g.vx, voxel_y=g.vy, voxel_z=g.vz, point_count=point_count(g),
        point_density=point_density(g, voxel_size), cog=cog, std=std_of_cog(x, y, z, g),
        cov_ref=covariance_reference_formula(x, y, z, g), cov_exact=covariance_exact(x, y, z, g),
        eig=eig, eig_n=eig_n, features=feat,
    )
