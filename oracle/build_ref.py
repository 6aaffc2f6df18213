"""Recipe: make the UNMODIFIED reference importable where /root/reference does not exist (the GPU box).

TEST / MEASUREMENT INFRASTRUCTURE.  The reference (3dgeo-heidelberg/vapc) is pure Python, so "building" it is copying its
package directory /root/reference/src/vapc, file for file and unchanged, to oracle/_ref/vapc — a directory that is git-ignored
(never in this repository's history) but not gpurun-ignored, so it travels to the GPU box like the built .so files.  There
`bench.py --impl reference` and the `cpu_baseline` leg time it through oracle/ref_harness.py (kind "reference"); without it they
fall back to oracle/ref_port.py (kind "port").  The reference's own test files (tests/test_voxel.py, tests/test_compute.py,
tests/integration_tests/test_main.py and test_io.py) and the three LAZ fixtures they read are copied the same way to oracle/_ref/tests for
tests/test_gpu_reference_suite.py.  Run by __graft_entry__.build() whenever /root/reference is present.

    python oracle/build_ref.py
"""
import filecmp
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.environ.get("VAPC_REFERENCE_SRC", "/root/reference/src")
DST = os.path.join(HERE, "_ref")
REFERENCE_UNIT_TESTS = ("test_voxel.py", "test_compute.py", "integration_tests/test_main.py", "integration_tests/test_io.py",
                        "test_data/vapc_in.laz", "test_data/small_mask.laz", "test_data/als_multichannel_leg000_points.laz",
                        "test_data/tree_wind_condition.laz")


def build(verbose=True) -> bool:
    src = os.path.join(SRC, "vapc")
    if not os.path.isdir(src):
        if verbose:
            print(f"oracle/build_ref.py: {src} not present; keeping oracle/_ref as it is")
        return os.path.isdir(os.path.join(DST, "vapc"))
    dst = os.path.join(DST, "vapc")
    os.makedirs(dst, exist_ok=True)
    for name in sorted(os.listdir(src)):
        if not name.endswith(".py"):
            continue
        a, b = os.path.join(src, name), os.path.join(dst, name)
        if not (os.path.exists(b) and filecmp.cmp(a, b, shallow=False)):
            shutil.copyfile(a, b)
    # the reference's own unit tests (pure Python, no fixtures on disk), unchanged as well: tests/test_gpu_reference_suite.py replays
    # them against the drop-in package (`import vapc` resolving to vapc_b200)
    tsrc, tdst = os.path.join(os.path.dirname(SRC), "tests"), os.path.join(DST, "tests")
    if os.path.isdir(tsrc):
        os.makedirs(tdst, exist_ok=True)
        for name in REFERENCE_UNIT_TESTS:
            a, b = os.path.join(tsrc, name), os.path.join(tdst, name)
            os.makedirs(os.path.dirname(b), exist_ok=True)
            if os.path.exists(a) and not (os.path.exists(b) and filecmp.cmp(a, b, shallow=False)):
                shutil.copyfile(a, b)
    if verbose:
        print(f"oracle/build_ref.py: reference package at {dst}")
    return True


if __name__ == "__main__":
    sys.exit(0 if build() else 1)
