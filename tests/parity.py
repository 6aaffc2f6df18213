"""Shared parity rules and golden-fixture helpers for the tests.

Parity rules, written once so that every test states the
tolerance it uses:

* integers (voxel triples, counts, mode, mode_count, mask membership): bit-exact.
* centroid, density, distances: |d| <= RTOL * |ref| (RTOL = 1e-9).
* std: |d| <= RTOL * |ref| + 8 * eps * max|coord|, the second term being the forward error of
  the reference's Welford recurrence (pandas group_var), whose running mean carries an absolute
  rounding error of eps*|coord| (measured: 1e-7 relative at UTM eastings on 1e-4 m spreads).
* covariance (H1): per voxel, max|dC| <= 1e-9 * max|C_ref| + 16 * eps * (Sxx+Syy+Szz)/n, the
  second term being the reference's own forward-error bound for its raw-moment formula.
* eigenvalues (H7): |d lambda_i| <= 1e-9 * lambda_1 per voxel.
* ratio features: absolute 1e-9; Sum_of_Eigenvalues relative 1e-9; Omnivariance through its
  cube, |d(l1 l2 l3)| <= 1e-9 * l1^3, reference-NaN accepted where l3_ref <= 1e-9 * l1.
* Mahalanobis d (bi_vapc.py:116-117): d = diff' inv(C + 1e-10 I) diff amplifies covariance
  differences by cond(C); rule |dd| <= d * (1e-9 + cov_rel * cond) where cov_rel is the relative
  covariance agreement that the H1 rule grants the pair being compared (1e-9 for GPU vs
  reference; 1e-12 for two evaluations of the same formula).  On well-conditioned voxels
  (cond <= 1e2) this is the plain 1e-9 rule.  p-values: abs 1e-15 + the same relative bound;
  significance / change_type exact away from the alpha threshold.
* closest-to-centroid (H5): the winner must lie in the reference's tied set
  {i : dist_ref(i) <= min_ref * (1 + 1e-12)}.
"""
import glob
import os

import numpy as np

RTOL = 1e-9
EPS = np.finfo(np.float64).eps
GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def golden_names():
    return sorted(os.path.basename(p)[:-4] for p in glob.glob(os.path.join(GOLDEN_DIR, "*.npz")))


def load_golden(name):
    """Load a fixture.  Reference quirk: on rank-deficient voxels LAPACK dgeev (np.linalg.eig,
    vapc.py:1060) sometimes returns the two noise eigenvalues as a complex-conjugate pair
    (|imag| ~ 1e-18), which turns the reference's whole Eigenvalue_* / feature columns into
    complex128.  The real parts are what is compared; imaginary parts are kept under
    `<key>__imag` and checked to be rounding noise by tests/test_oracle_golden.py."""
    out = {}
    for k, a in np.load(os.path.join(GOLDEN_DIR, name + ".npz"), allow_pickle=False).items():
        if a.dtype.kind == "c":
            out[k + "__imag"] = a.imag.copy()
            a = a.real.copy()
        out[k] = a
    return out


def cloud_cases():
    return [n for n in golden_names() if n not in ("kat_voxelize", "clusters_nn", "label_change")]


def per_voxel(gold, point2voxel, nvox, col):
    """Collapse a per-point golden column to one value per voxel (checks it is constant)."""
    a = gold["pt_" + col]
    out = np.empty(nvox, a.dtype)
    out[point2voxel] = a
    back = out[point2voxel]
    same = (back == a) | (np.isnan(back.astype(float)) & np.isnan(a.astype(float)))
    if col == "Omnivariance":  # complex principal cube roots of noise products (see load_golden)
        same |= np.abs(gold.get("pt_Omnivariance__imag", np.zeros(a.shape))) > 0
    assert same.all(), f"{col} is not constant per voxel in the golden"
    return out


def assert_close_rel(got, ref, rtol=RTOL, what="", atol=0.0):
    got = np.asarray(got, np.float64)
    ref = np.asarray(ref, np.float64)
    assert got.shape == ref.shape, (what, got.shape, ref.shape)
    nan_g, nan_r = np.isnan(got), np.isnan(ref)
    assert (nan_g == nan_r).all(), f"{what}: NaN pattern differs ({nan_g.sum()} vs {nan_r.sum()})"
    ok = ~nan_r
    err = np.abs(got[ok] - ref[ok])
    bound = rtol * np.abs(ref[ok]) + atol
    bad = err > bound
    assert not bad.any(), f"{what}: {bad.sum()} values exceed rtol {rtol} atol {atol}; worst rel {np.max(err / np.maximum(np.abs(ref[ok]), 1e-300)):.3e}"


def assert_cov_h1(got6, ref6, counts, sumsq_over_n, what="cov"):
    """H1 rule.  got6/ref6: [V, 6]; sumsq_over_n: [V] = (Sxx+Syy+Szz)/n of raw coordinates."""
    nan_g, nan_r = np.isnan(got6).any(axis=1), np.isnan(ref6).any(axis=1)
    assert (nan_g == nan_r).all(), f"{what}: NaN rows differ"
    assert (nan_r == (counts <= 1)).all()
    ok = ~nan_r
    err = np.max(np.abs(got6[ok] - ref6[ok]), axis=1)
    bound = RTOL * np.max(np.abs(ref6[ok]), axis=1) + 16 * EPS * sumsq_over_n[ok]
    bad = err > bound
    assert not bad.any(), f"{what}: {bad.sum()} voxels exceed the H1 bound; worst ratio {np.max(err / bound):.3e}"


def noise_floor(cmax):
    """Absolute size of the rounding noise in a covariance / eigenvalue that the reference computes by
    centring with an fp64 mean (np.cov): the mean is off by ~eps*|coord|, its square is the floor."""
    return (8 * EPS * cmax) ** 2


def assert_eig_h7(got3, ref3, counts, what="eig", abs_floor=0.0):
    nan_r = np.isnan(ref3).any(axis=1)
    assert (np.isnan(got3).any(axis=1) == nan_r).all(), f"{what}: NaN rows differ"
    assert (nan_r == (counts <= 1)).all()
    ok = ~nan_r
    l1 = ref3[ok, 0:1]
    err = np.abs(got3[ok] - ref3[ok])
    bad = err > RTOL * l1 + abs_floor
    assert not bad.any(), f"{what}: {bad.sum()} eigenvalues off; worst {np.max(err / np.maximum(l1, 1e-300)):.3e} of lambda_1"


FEATURES = ("Sum_of_Eigenvalues", "Omnivariance", "Eigenentropy", "Anisotropy", "Planarity", "Linearity",
            "Surface_Variation", "Sphericity")


def assert_features_h7(got8, ref8, ref_eig, counts, what="features", abs_floor=0.0):
    n1 = counts <= 1
    assert np.isnan(got8[n1]).all() and np.isnan(ref8[n1]).all(), f"{what}: n=1 voxels must be NaN"
    # voxels whose largest reference eigenvalue is itself rounding noise (all points identical): every
    # feature of the reference is noise / noise there, nothing to compare
    with np.errstate(invalid="ignore"):
        ok = ~n1 & (ref_eig[:, 0] > 1e9 * abs_floor)
    g, r, e = got8[ok], ref8[ok], ref_eig[ok]
    l1 = e[:, 0]
    # Sum: relative
    assert (np.abs(g[:, 0] - r[:, 0]) <= RTOL * np.abs(r[:, 0])).all(), f"{what}: Sum_of_Eigenvalues"
    # Omnivariance through its cube; reference NaN accepted on rank-deficient voxels
    degenerate = e[:, 2] <= RTOL * l1
    cube_g, cube_r = g[:, 1] ** 3, r[:, 1] ** 3
    with np.errstate(invalid="ignore"):
        bad = ~(np.abs(cube_g - cube_r) <= RTOL * l1 ** 3)
    bad &= ~degenerate  # rank-deficient voxels: the reference value is noise^(1/3), NaN or complex
    assert not bad.any(), f"{what}: Omnivariance ({bad.sum()} voxels)"
    assert not np.isnan(g[~degenerate, 1]).any(), f"{what}: Omnivariance NaN on a full-rank voxel"
    for k in range(2, 8):
        err = np.abs(g[:, k] - r[:, k])
        assert (err <= RTOL).all(), f"{what}: {FEATURES[k]} worst abs err {np.nanmax(err):.3e}"


def assert_mahalanobis(got, ref, cov9_plus_eps, cov_rel, what="d"):
    """got/ref: [R]; cov9_plus_eps: [R, 3, 3] the matrices that were inverted; cov_rel: scalar or [R]
    relative covariance agreement granted to the pair being compared."""
    nan_r = np.isnan(ref)
    assert (np.isnan(got) == nan_r).all(), f"{what}: NaN pattern"
    ok = ~nan_r
    cond = np.linalg.cond(cov9_plus_eps[ok])
    rel = np.broadcast_to(np.asarray(cov_rel, np.float64), ref.shape)[ok]
    err = np.abs(got[ok] - ref[ok])
    bound = np.abs(ref[ok]) * (RTOL + rel * cond)
    bad = err > bound
    assert not bad.any(), f"{what}: {bad.sum()} rows exceed the conditioned bound; worst {np.max(err / bound):.3e}"
    return cond


def h1_cov_rel(cov9, sumsq_over_n):
    """Relative covariance agreement the H1 rule grants against the reference's raw-moment formula:
    1e-9 + 16 eps (Sxx+Syy+Szz)/n / max|C|."""
    cmax = np.nanmax(np.abs(cov9.reshape(len(cov9), -1)), axis=1)
    with np.errstate(divide="ignore", invalid="ignore"):
        return RTOL + 16 * EPS * sumsq_over_n / cmax
