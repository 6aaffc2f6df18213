"""The reference's OWN unit tests, unmodified, run against the drop-in package.

oracle/build_ref.py copies /root/reference/tests/test_voxel.py and test_compute.py unchanged into the git-ignored oracle/_ref/tests
(they travel to the GPU box with the snapshot, like oracle/_ref/vapc).  Here they run in a pytest subprocess in which `import vapc`
resolves to vapc_b200 (a three-line shim module on PYTHONPATH; oracle/_ref/vapc is NOT importable there), so every assertion the
reference's authors wrote about voxelize / voxel index / corners / buffer / masks / the three reductions / offsets / clusters / the
attribute dispatcher (mock.patch of `vapc.Vapc.compute_*`) / per-attribute statistics incl. `mode` of a negative, continuous
attribute / filter_attributes is made about this implementation.
"""
import os
import re
import shutil
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_TESTS = os.path.join(ROOT, "oracle", "_ref", "tests")
FILES = ("test_voxel.py", "test_compute.py")


@pytest.mark.gpu
def test_reference_unit_tests_pass_against_the_drop_in(tmp_path):
    if not all(os.path.exists(os.path.join(REF_TESTS, f)) for f in FILES):
        pytest.skip("oracle/_ref/tests not built (oracle/build_ref.py needs /root/reference)")
    shim = tmp_path / "shim"
    shim.mkdir()
    (shim / "vapc.py").write_text("import sys\n\nimport vapc_b200\n\nsys.modules['vapc'] = vapc_b200\n")
    run_dir = tmp_path / "reference_tests"
    run_dir.mkdir()
    for f in FILES:
        shutil.copyfile(os.path.join(REF_TESTS, f), run_dir / f)  # byte-identical copies: __pycache__ stays out of oracle/_ref
    (run_dir / "pytest.ini").write_text("[pytest]\n")
    env = dict(os.environ, PYTHONPATH=os.pathsep.join([str(shim), ROOT]))
    r = subprocess.run([sys.executable, "-m", "pytest", "-q", "-p", "no:cacheprovider", "-c", str(run_dir / "pytest.ini"), "--rootdir", str(run_dir),
                        *[str(run_dir / f) for f in FILES]], env=env, cwd=str(run_dir), capture_output=True, text=True, timeout=900)
    tail = r.stdout[-6000:] + r.stderr[-2000:]
    print(tail)
    assert r.returncode == 0, tail
    summary = r.stdout.strip().splitlines()[-1]
    passed = re.search(r"(\d+) passed", summary)
    assert passed and int(passed.group(1)) >= 31 and "failed" not in summary and "error" not in summary, summary
