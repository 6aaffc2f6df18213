"""CPU tests (no GPU): the C-ABI library loads and exports every symbol include/vapc_b200.h
declares, the ctypes table matches the header, host-side helpers and the shared fp64 math (compiled
for the host as self-test entry points) agree with numpy, and the facade validates its arguments
like the reference.  No GPU compute is called here."""
import ctypes as C
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_functions():
    src = open(os.path.join(ROOT, "include", "vapc_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    out = {}
    for m in re.finditer(r"VAPC_API\s+[\w\s\*]+?\b(vapc_\w+)\s*\(([^;]*?)\)\s*;", src, flags=re.S):
        args = m.group(2).strip()
        out[m.group(1)] = 0 if args in ("", "void") else args.count(",") + 1
    return out


def test_library_exports_every_declared_symbol():
    from vapc_b200 import _lib as L

    decl = header_functions()
    assert len(decl) >= 35
    lib = C.CDLL(L.LIB_PATH)
    for name in decl:
        assert hasattr(lib, name), f"{name} declared in include/vapc_b200.h but not exported"
    # the ctypes table binds exactly the declared functions, with the declared number of arguments
    assert set(L.SIGNATURES) == set(decl)
    for name, (_, argtypes) in L.SIGNATURES.items():
        assert len(argtypes) == decl[name], name
    assert L.raw("vapc_abi_version")() == 1


def test_no_cpu_fallback_without_gpu():
    import torch

    from vapc_b200 import engine as E

    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        E.VoxelCloud.build(np.zeros(4), np.zeros(4), np.zeros(4), 1.0)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        E.voxel_coords(np.zeros(4), np.zeros(4), np.zeros(4), 1.0, [0, 0, 0])


def test_product_never_imports_oracle():
    for dirpath, _, files in os.walk(os.path.join(ROOT, "vapc_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".sh")):
                text = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in text.replace("no oracle", ""), f"{f} mentions the oracle"


def test_grid_from_bounds_matches_numpy_floor():
    from vapc_b200 import engine as E
    from vapc_b200._lib import VapcError

    rng = np.random.default_rng(0)
    for _ in range(200):
        vs = float(rng.choice([0.05, 0.1, 0.25, 0.3, 1.0, 7.5]))
        origin = rng.uniform(-10, 10, 3).tolist()
        lo = rng.uniform(-1e4, 1e4, 3)
        hi = lo + rng.uniform(0, 500, 3)
        g = E.grid_from_bounds(vs, origin, list(lo) + list(hi))
        vmin = np.floor((lo - origin) / vs).astype(np.int64)
        vmax = np.floor((hi - origin) / vs).astype(np.int64)
        assert list(g.vmin) == vmin.tolist()
        for k in range(3):
            assert g.bits[k] == int(vmax[k] - vmin[k]).bit_length()
        assert g.key_bytes == (4 if g.total_bits <= 32 else 8)
    g = E.grid_from_bounds(0.1, [0, 0, 0], [0.3, 0.6, 0.7, 0.3, 0.6, 0.7])  # KAT-1: 2, 5, 6 (not 3, 6, 7)
    assert list(g.vmin) == [2, 5, 6] and g.total_bits == 0
    with pytest.raises(VapcError):
        E.grid_from_bounds(1e-9, [0, 0, 0], [0, 0, 0, 1e6, 1e6, 1e6])  # 3 x 50 bits
    with pytest.raises(ValueError):
        E.grid_from_bounds(-1.0, [0, 0, 0], [0, 0, 0, 1, 1, 1])
    with pytest.raises(ValueError):
        E.grid_from_bounds(1.0, [0, 0, 0], [0, 0, 0, np.nan, 1, 1])


def test_host_selftest_eig3_matches_numpy():
    from vapc_b200 import _lib as L

    rng = np.random.default_rng(1)
    mats = []
    for i in range(4000):
        n = int(rng.integers(2, 12))
        pts = rng.normal(0, 1, (n, 3)) * rng.uniform(1e-4, 1, 3)
        if i % 7 == 0:
            pts[:, 2] = pts[:, 0] * 0.5  # rank deficient
        if i % 11 == 0:
            pts = np.round(pts, 2)
        mats.append(np.cov(pts.T))
    mats += [np.zeros((3, 3)), np.eye(3), np.diag([3.0, 3.0, 1.0]), np.diag([1e-12, 5.0, 5.0]), np.full((3, 3), 2.0)]
    mats = np.array(mats)
    cov6 = np.ascontiguousarray(mats[:, [0, 0, 0, 1, 1, 2], [0, 1, 2, 1, 2, 2]])
    out = np.empty((len(mats), 3))
    assert L.raw("vapc_selftest_eig3_host")(cov6.ctypes.data, len(mats), out.ctypes.data) == 0
    ref = np.linalg.eigvalsh(mats)[:, ::-1]
    scale = np.maximum(np.abs(ref[:, :1]), 1e-300)
    assert np.max(np.abs(out - ref) / scale) < 1e-13
    assert np.all(out[:, 0] >= out[:, 1]) and np.all(out[:, 1] >= out[:, 2])


def test_host_selftest_chan_combination():
    from vapc_b200 import _lib as L

    rng = np.random.default_rng(2)
    pts = rng.normal(0, 0.02, (5000, 3)) + [0.21, -0.13, 0.4]
    cuts = [0, 1, 2, 40, 41, 1000, 1001, 4999, 5000]
    rows = []
    for a, b in zip(cuts[:-1], cuts[1:]):
        c = pts[a:b]
        m = c.mean(0)
        d = c - m
        M = d.T @ d
        rows.append([b - a, *m, M[0, 0], M[0, 1], M[0, 2], M[1, 1], M[1, 2], M[2, 2]])
    rows = np.ascontiguousarray(np.array(rows, dtype=np.float64))
    out = np.empty(10)
    assert L.raw("vapc_selftest_chan_host")(rows.ctypes.data, len(rows), out.ctypes.data) == 0
    m = pts.mean(0)
    d = pts - m
    M = d.T @ d
    ref = np.array([5000, *m, M[0, 0], M[0, 1], M[0, 2], M[1, 1], M[1, 2], M[2, 2]])
    np.testing.assert_allclose(out, ref, rtol=1e-12, atol=1e-18)


def test_host_selftest_mahalanobis_matches_scipy():
    from scipy.stats import chi2

    from vapc_b200 import _lib as L

    rng = np.random.default_rng(3)
    n = 3000
    covs, diffs = [], []
    for i in range(n):
        npts = 2 if i % 5 == 0 else (3 if i % 5 == 1 else 8)  # rank-1 / rank-2 covariances: cond ~ 1e8 after + 1e-10 I
        pts = rng.normal(0, 1, (npts, 3)) * rng.uniform(0.01, 0.3, 3)
        covs.append(np.cov(pts.T))
        diffs.append(rng.normal(0, 0.2, 3) * rng.choice([0.01, 0.1, 1.0, 10.0]))
    covs, diffs = np.array(covs), np.ascontiguousarray(np.array(diffs))
    cov6 = np.ascontiguousarray(covs[:, [0, 0, 0, 1, 1, 2], [0, 1, 2, 1, 2, 2]])
    d = np.empty(n)
    p = np.empty(n)
    assert L.raw("vapc_selftest_mahalanobis_host")(cov6.ctypes.data, diffs.ctypes.data, n, d.ctypes.data, p.ctypes.data) == 0
    ce = covs + np.eye(3) * 1e-10
    d_ref = np.einsum("ij,ijk,ik->i", diffs, np.linalg.inv(ce), diffs)
    cond = np.linalg.cond(ce)
    assert cond.max() > 1e7
    assert np.all(np.abs(d - d_ref) <= d_ref * (1e-12 + 1e-14 * cond))  # both are backward-stable pivoted solves
    p_ref = 1 - chi2.cdf(d_ref, df=3)
    well = cond < 1e3
    assert np.all(np.abs(p - p_ref)[well] <= 1e-15 + 1e-9 * p_ref[well])
    # the tail: p is exactly 0 where scipy's 1 - cdf is (drives change_type 2, bi_vapc.py:135)
    big = np.array([60.0, 75.0, 78.0, 79.0, 80.0, 100.0, 1e4, 1e12])
    cov6 = np.tile(np.array([1.0 - 1e-10, 0, 0, 1.0 - 1e-10, 0, 1.0 - 1e-10]), (len(big), 1))
    diffs = np.zeros((len(big), 3))
    diffs[:, 0] = np.sqrt(big)
    d = np.empty(len(big))
    p = np.empty(len(big))
    L.raw("vapc_selftest_mahalanobis_host")(np.ascontiguousarray(cov6).ctypes.data, np.ascontiguousarray(diffs).ctypes.data, len(big),
                                            d.ctypes.data, p.ctypes.data)
    p_ref = 1 - chi2.cdf(d, df=3)
    assert np.array_equal(p == 0, p_ref == 0)
    assert np.all(np.abs(p - p_ref) <= 4e-16)


def test_facade_argument_validation_matches_reference():
    import pandas as pd

    import vapc_b200 as vapc

    with pytest.raises(ValueError):
        vapc.Vapc(0)
    with pytest.raises(ValueError):
        vapc.Vapc(-1.0)
    with pytest.raises(ValueError):
        vapc.Vapc(1.0, origin=[0, 0])
    with pytest.raises(ValueError):
        vapc.Vapc(1.0, origin=(0, 0, 0))  # the reference insists on a list (vapc.py:68)
    v = vapc.Vapc(0.5)
    assert v.origin == [0, 0, 0] and v.return_at == "closest_to_center_of_gravity" and v.compute == [] and v.attributes == {}
    assert len(v.AVAILABLE_COMPUTATIONS) == 13
    v.compute = ["point_count", "nonsense"]
    with pytest.raises(ValueError):
        v.compute_requested_attributes()
    v.df = pd.DataFrame({"X": [0.0, 1.0], "Y": [0.0, 1.0], "Z": [0.0, 1.0], "point_count": [1, 5]})
    with pytest.raises(ValueError):
        v.filter_attributes("point_count", "leq", 3)
    with pytest.raises(KeyError):
        v.filter_attributes("attribute", "<=", 3)
    v.filter_attributes("point_count", "Greater_Than", 3)  # case-insensitive long names
    assert len(v.df) == 1
    with pytest.raises(ValueError):
        vapc.BiTemporalVapc([v])
    with pytest.raises(AttributeError):
        class DH:
            df = None
        vapc.Vapc(1.0).get_data_from_data_handler(DH())
    assert callable(vapc.enable_trace) and callable(vapc.enable_timeit)


def test_histogram_mode_restatement_matches_reference():
    """vapc_b200.utilities.compute_mode_continuous (the fallback of `mode` for values np.bincount rejects) against the unmodified
    reference's function (utilities.py:114-159), values and exceptions."""
    from oracle import ref_harness

    if not ref_harness.available():
        pytest.skip("reference package not present")
    ref_harness.import_reference()
    from vapc.utilities import compute_mode_continuous as ref_mode

    from vapc_b200.utilities import compute_mode_continuous as ours

    rng = np.random.default_rng(0)
    for _ in range(500):
        d = np.round(rng.normal(-5, 4, rng.integers(5, 300)), 1)
        assert ours(d) == ref_mode(d)
    for bad in (np.array([-1.0]), np.array([-2.0, -2.0, -2.0]), np.array([-1.0, np.nan, 3.0])):
        with pytest.raises(ValueError):
            ref_mode(bad)
        with pytest.raises(ValueError):
            ours(bad)


def test_segmented_histogram_modes_equal_numpy_bit_for_bit():
    """engine.segmented_histogram_modes (the batch form of the `mode` fallback, tensor ops only: the same code runs on the GPU)
    against the per-sample numpy function — and through it the reference's (test above) — on samples of every kind the statistic
    meets: rounded scan angles, small integers with ties between bins, wide dynamic range, values at UTM magnitude; batches holding a
    sample the reference raises for are refused as a whole (the caller then takes the per-sample route and raises like numpy)."""
    import torch

    from vapc_b200 import engine as E
    from vapc_b200.utilities import compute_mode_continuous

    rng = np.random.default_rng(3)
    compared = 0
    for _ in range(12):
        samples = []
        for _s in range(int(rng.integers(1, 300))):
            n = int(rng.integers(2, 200)) if rng.random() < 0.9 else int(rng.integers(200, 3000))
            kind = rng.integers(0, 5)
            if kind == 0:
                v = np.round(rng.normal(-5, 4, n), 1)
            elif kind == 1:
                v = rng.integers(-30, 30, n).astype(np.float64)
            elif kind == 2:
                v = rng.normal(0, 1, n) * 10.0 ** rng.integers(-6, 6)
            elif kind == 3:
                v = np.round(rng.uniform(-90, 90, n))
            else:
                v = rng.integers(-3, 3, n) * 0.1 + 476850.0
            samples.append(v)
        ref, fine = [], []
        for v in samples:
            try:
                with np.errstate(all="ignore"):
                    ref.append(compute_mode_continuous(v))
                fine.append(v)
            except (ValueError, OverflowError):
                pass
        if fine:
            got = E.segmented_histogram_modes(torch.from_numpy(np.concatenate(fine)), torch.tensor([len(v) for v in fine]))
            np.testing.assert_array_equal(got.numpy(), np.array(ref))
            compared += len(ref)
        if len(fine) != len(samples):
            assert E.segmented_histogram_modes(torch.from_numpy(np.concatenate(samples)), torch.tensor([len(v) for v in samples])) is None
    assert compared > 1000
    assert E.segmented_histogram_modes(torch.zeros(0, dtype=torch.float64), torch.zeros(0, dtype=torch.int64)).numel() == 0
    assert E.segmented_histogram_modes(torch.tensor([-1.0, float("nan"), 3.0], dtype=torch.float64), torch.tensor([3])) is None
    assert E.segmented_histogram_modes(torch.tensor([-2.0, -2.0, -2.0, 1.0, 5.0], dtype=torch.float64), torch.tensor([3, 2])) is None


def test_cloud_signature_sees_column_edits():
    """Vapc._signature (what the cached device cloud is compared with): same frame -> same signature; a shifted, replaced or
    re-assigned coordinate column, another voxel size or origin -> another one (host logic only, no GPU needed)."""
    import pandas as pd

    import vapc_b200 as vapc

    rng = np.random.default_rng(5)
    df = pd.DataFrame({"X": rng.uniform(0, 9, 5000), "Y": rng.uniform(0, 9, 5000), "Z": rng.uniform(0, 3, 5000), "intensity": 1.0})
    v = vapc.Vapc(0.5)
    v.df = df
    s0 = v._signature()
    assert v._signature() == s0
    df["intensity"] = 2.0
    assert v._signature() == s0  # not a coordinate
    df["Y"] = df["Y"].to_numpy() * 2.0
    s1 = v._signature()
    assert s1 != s0
    df.loc[:, "Z"] = df["Z"].to_numpy() + 1.0
    assert v._signature() != s1
    v.voxel_size = 0.25
    s2 = v._signature()
    v.origin = [1.0, 0.0, 0.0]
    assert len({s0, s1, s2, v._signature()}) == 4
    e = vapc.Vapc(0.5)
    e.df = df.iloc[:0]
    assert e._signature() == e._signature()
    e.invalidate()  # nothing computed yet: a no-op


def test_header_is_plain_c_and_links_from_a_c_program(tmp_path):
    """The boundary is a C ABI: include/vapc_b200.h compiles as strict C99 (no C++, no torch types), and a C program linked against
    libvapc_b200.so alone calls it — the host-only entry points here (version, workspace sizes, the grid from a bounding box, the
    argument check of a compute entry point), no GPU needed."""
    import shutil
    import subprocess

    from vapc_b200 import _lib

    if shutil.which("gcc") is None:
        pytest.skip("no gcc")
    src = tmp_path / "client.c"
    src.write_text(r'''
#include <stdio.h>
#include <string.h>
#include "vapc_b200.h"
int main(void) {
  vapc_grid_t g;
  double mm[6] = {0.0, 0.0, 0.0, 49.9999, 49.9999, 3.9999};
  memset(&g, 0, sizeof g);
  g.voxel_size = 0.1;
  if (vapc_abi_version() <= 0) return 1;
  if (vapc_sort_workspace_bytes(100000000) == 0 || vapc_segment_workspace_bytes(100000000) == 0) return 2;
  if (vapc_grid_from_bounds(&g, mm) != 0) return 3;
  if (g.bits[0] != 9 || g.bits[1] != 9 || g.bits[2] != 6 || g.key_bytes != 4) return 4; /* 500 x 500 x 40 voxels: 24 key bits */
  if (vapc_cog_reference_formula(NULL, NULL, NULL, -1, NULL, NULL) == 0) return 5;       /* bad arguments are refused ... */
  if (strlen(vapc_last_error()) == 0) return 6;                                          /* ... with a message */
  printf("abi %d bits %d %d %d\n", vapc_abi_version(), g.bits[0], g.bits[1], g.bits[2]);
  return 0;
}
''')
    inc = os.path.join(ROOT, "include")
    subprocess.run(["gcc", "-std=c99", "-Wall", "-Wextra", "-Werror", "-pedantic", "-I", inc, "-fsyntax-only", str(src)], check=True)
    exe = tmp_path / "client"
    libdir = os.path.dirname(_lib.LIB_PATH)
    subprocess.run(["gcc", "-std=c99", "-I", inc, str(src), "-o", str(exe), "-L", libdir, "-l:" + os.path.basename(_lib.LIB_PATH),
                    "-Wl,-rpath," + libdir], check=True)
    out = subprocess.run([str(exe)], capture_output=True, text=True)
    assert out.returncode == 0, (out.returncode, out.stdout, out.stderr)
    assert out.stdout.startswith("abi ")


def test_numpy_sum_order_of_the_attribute_kernel_on_the_host(tmp_path):
    """csrc/np_sum.cuh — the summation order attr_stats_kernel uses for mean / sum / var — compiled for the host from the same text
    (plain IEEE operations, -ffp-contract=off) equals ndarray.sum / mean / var bit for bit on slices of 1 to 1e6 values, contiguous
    and as the strided field of a structured array (what the reference slices, vapc.py:387-399)."""
    import ctypes
    import shutil
    import subprocess

    if shutil.which("g++") is None:
        pytest.skip("no g++")
    src = tmp_path / "np_sum_host.cpp"
    src.write_text('''
#include "np_sum.cuh"
extern "C" void stats(const double* x, long long stride, long long n, double* out) {
  auto value = [&](int64_t p) { return x[p * stride]; };
  const double s = np_pairwise_sum(value, 0, n);
  const double m = s / (double)n;
  auto sq = [&](int64_t p) { const double d = value(p) - m; return d * d; };
  out[0] = s; out[1] = m; out[2] = np_pairwise_sum(sq, 0, n) / (double)n;
}
''')
    so = tmp_path / "np_sum_host.so"
    subprocess.run(["g++", "-O2", "-ffp-contract=off", "-shared", "-fPIC", "-I", os.path.join(ROOT, "vapc_b200", "csrc"), "-o", str(so), str(src)], check=True)
    lib = ctypes.CDLL(str(so))
    rng = np.random.default_rng(5)
    sizes = list(range(1, 300)) + [511, 512, 513, 1000, 1024, 4097, 8191, 8192, 8193, 10000, 65536, 100003, 1000000]
    out = (ctypes.c_double * 3)()
    for n in sizes:
        a = rng.normal(-3, 5, n) * 10.0 ** rng.integers(-3, 6)
        rec = np.zeros(n, dtype=[("voxel_x", "i8"), ("voxel_y", "i8"), ("voxel_z", "i8"), ("attr", "f8")])
        rec["attr"] = a
        for arr in (a, rec["attr"]):
            lib.stats(ctypes.c_void_p(arr.ctypes.data), ctypes.c_longlong(arr.strides[0] // 8), ctypes.c_longlong(n), out)
            assert tuple(out) == (float(arr.sum()), float(arr.mean()), float(arr.var())), n


def test_module_level_helpers_match_reference_on_the_host():
    """utilities.optimal_bins / get_version and bi_vapc.merge_with_suffix (names notebooks import from the reference) against the
    unmodified reference's functions; compute_euclidean_distance / mutual_nearest_neighbors need the GPU (tests/test_gpu_neighbors.py)
    and refuse to run without one."""
    import pandas as pd

    import vapc_b200
    from vapc_b200 import bi_vapc as B
    from vapc_b200 import utilities as U

    assert U.get_version() == vapc_b200.get_version()
    a, b = pd.DataFrame({"cog_x": [1.0, 2.0], "n": [3, 4]}), pd.DataFrame({"cog_x": [5.0, 6.0], "n": [7, 8]})
    m = B.merge_with_suffix(a, b)
    assert m.columns.tolist() == ["cog_x_x", "n_x", "cog_x_y", "n_y"] and m["n_y"].tolist() == [7, 8]
    with pytest.raises(ValueError):
        U.optimal_bins(np.arange(10.0), "doane")
    with pytest.raises(RuntimeError):
        B.compute_euclidean_distance(pd.DataFrame({f"cog_{c}_{s}": [0.0] for c in "xyz" for s in "xy"}))
    from oracle import ref_harness

    if not ref_harness.available():
        pytest.skip("reference package not present")
    ref_harness.import_reference()
    from vapc.utilities import optimal_bins as ref_bins

    rng = np.random.default_rng(2)
    for n in (10, 100, 1000):
        d = rng.normal(0, 3, n)
        for method in ("sqrt", "sturges", "rice", "fd", "scott"):
            assert U.optimal_bins(d, method) == ref_bins(d, method), (n, method)
