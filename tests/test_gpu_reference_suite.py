"""The reference's OWN tests, unmodified, run against the drop-in package.

oracle/build_ref.py copies the reference's tests/test_voxel.py, tests/test_compute.py, tests/integration_tests/test_main.py and
test_io.py and the three LAZ fixtures they read, unchanged, into the git-ignored oracle/_ref/tests (they travel to the GPU box with the
snapshot, like oracle/_ref/vapc).  Here they run in a pytest subprocess whose conftest.py — the only file of ours in that tree — imports
vapc_b200 and registers it as the module `vapc` (oracle/_ref/vapc is NOT importable there), so every assertion the reference's authors
wrote is made about this implementation:

* unit tests: voxelize / voxel index / corners / buffer / masks / the three reductions / offsets / clusters / the attribute dispatcher
  (mock.patch of `vapc.Vapc.compute_*`) / per-attribute statistics incl. `mode` of a negative, continuous attribute / filter_attributes;
* integration tests: DataHandler on the reference's LAZ files (shapes, attribute names, error types), LAS 1.4 / PLY writing,
  do_vapc_on_files file to file (`13.43 percent of the voxel space is occupied`, 13 522 voxels of the two merged files, .las / .laz / .ply).

The integration tests inspect the written files with `laspy` and `plyfile`, which this image does not have: tests/ref_shims holds a
reader of each (test infrastructure over vapc_b200.las / numpy) that becomes importable only after vapc_b200 itself has been imported.
"""
import os
import re
import shutil
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_TESTS = os.path.join(ROOT, "oracle", "_ref", "tests")
SHIMS = os.path.join(ROOT, "tests", "ref_shims")
CONFTEST = f'''import sys

import vapc_b200  # first: the package under test must not see the laspy / plyfile stand-ins of the tests

sys.modules["vapc"] = vapc_b200
sys.path.insert(0, {SHIMS!r})
'''


def replay(tmp_path, files, data=()):
    if not all(os.path.exists(os.path.join(REF_TESTS, f)) for f in tuple(files) + tuple(data)):
        pytest.skip("oracle/_ref/tests not built (oracle/build_ref.py needs /root/reference)")
    run_dir = tmp_path / "reference"
    for f in tuple(files) + tuple(data):  # byte-identical copies: __pycache__ and outputs stay out of oracle/_ref
        os.makedirs(run_dir / "tests" / os.path.dirname(f), exist_ok=True)
        shutil.copyfile(os.path.join(REF_TESTS, f), run_dir / "tests" / f)
    (run_dir / "pytest.ini").write_text("[pytest]\n")
    (run_dir / "conftest.py").write_text(CONFTEST)
    env = dict(os.environ, PYTHONPATH=ROOT)
    r = subprocess.run([sys.executable, "-m", "pytest", "-q", "-p", "no:cacheprovider", "-c", str(run_dir / "pytest.ini"), "--rootdir", str(run_dir),
                        *[str(run_dir / "tests" / f) for f in files]], env=env, cwd=str(run_dir), capture_output=True, text=True, timeout=1500)
    tail = r.stdout[-8000:] + r.stderr[-2000:]
    print(tail)
    summary = r.stdout.strip().splitlines()[-1] if r.stdout.strip() else ""
    passed = re.search(r"(\d+) passed", summary)
    return r.returncode, int(passed.group(1)) if passed else 0, summary, tail


@pytest.mark.gpu
def test_reference_unit_tests_pass_against_the_drop_in(tmp_path):
    rc, passed, summary, tail = replay(tmp_path, ("test_voxel.py", "test_compute.py"))
    assert rc == 0 and passed >= 31 and "failed" not in summary and "error" not in summary, tail


@pytest.mark.gpu
def test_reference_integration_tests_pass_against_the_drop_in(tmp_path):
    rc, passed, summary, tail = replay(tmp_path, ("integration_tests/test_main.py", "integration_tests/test_io.py"),
                                       data=("test_data/vapc_in.laz", "test_data/small_mask.laz", "test_data/als_multichannel_leg000_points.laz"))
    assert rc == 0 and passed >= 30 and "failed" not in summary and "error" not in summary, tail
