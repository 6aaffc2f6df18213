"""Import harness for the UNMODIFIED reference (3dgeo-heidelberg/vapc) in THIS container.

TEST / MEASUREMENT INFRASTRUCTURE ONLY.  Used by `oracle/make_golden.py` to generate the committed
fixtures under `tests/golden/` (from `/root/reference`, in the build container) and by `bench.py`'s CPU
legs to TIME the unmodified reference on the GPU box's host cores (from `oracle/_ref`, the unchanged copy of the
reference package that `oracle/build_ref.py` makes; `/root/reference` does not exist there).

The reference cannot be imported as-is here because `src/vapc/__init__.py:3,5-7` pulls in
`datahandler`/`main`, which need `laspy` and `plyfile` (not installed, no network).  We
register empty stub modules for those two (the hot path never touches them) and, because
`bi_vapc.py:109` mutates the array returned by `DataFrame.to_numpy()` (read-only under
pandas-3 copy-on-write), make `to_numpy` return a writable copy.  The reference tree itself
is not modified.
"""
import os
import sys
import types

_HERE = os.path.dirname(os.path.abspath(__file__))


def _find_reference():
    for cand in (os.environ.get("VAPC_REFERENCE_SRC"), "/root/reference/src", os.path.join(_HERE, "_ref")):
        if cand and os.path.isdir(os.path.join(cand, "vapc")):
            return cand
    return "/root/reference/src"


REFERENCE_SRC = _find_reference()


def available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_SRC, "vapc"))


def full_attribute_set(x, y, z, voxel_size, origin=(0.0, 0.0, 0.0)):
    """BASELINE configs[1] through the reference's own public API (Vapc + compute list, vapc.py:211-252): voxelize, point_count,
    point_density, center_of_gravity, std_of_cog, covariance_matrix, eigenvalues, geometric_features.  Returns the Vapc object."""
    import warnings

    import pandas as pd

    vapc = import_reference()
    warnings.filterwarnings("ignore")
    v = vapc.Vapc(voxel_size, list(origin))
    v.df = pd.DataFrame({"X": x, "Y": y, "Z": z})
    v.original_attributes = list(v.df.columns)
    v.compute = ["point_count", "point_density", "center_of_gravity", "std_of_cog", "covariance_matrix", "eigenvalues", "geometric_features"]
    v.compute_requested_attributes()
    return v


def import_reference():
    """Return the reference `vapc` module (raises if /root/reference is absent)."""
    if not os.path.isdir(os.path.join(REFERENCE_SRC, "vapc")):
        raise RuntimeError(f"reference sources not found under {REFERENCE_SRC}")
    if "vapc" in sys.modules and getattr(sys.modules["vapc"], "__file__", "").startswith(REFERENCE_SRC):
        return sys.modules["vapc"]
    if "laspy" not in sys.modules:
        laspy = types.ModuleType("laspy")
        laspy.errors = types.ModuleType("laspy.errors")

        class LaspyException(Exception):
            pass

        laspy.errors.LaspyException = laspy.LaspyException = LaspyException
        sys.modules["laspy"] = laspy
        sys.modules["laspy.errors"] = laspy.errors
    if "plyfile" not in sys.modules:
        ply = types.ModuleType("plyfile")
        ply.PlyData = ply.PlyElement = object
        sys.modules["plyfile"] = ply
    import pandas as pd

    if not getattr(pd.DataFrame.to_numpy, "_vapc_writable", False):
        _tn = pd.DataFrame.to_numpy

        def to_numpy(self, *a, **k):
            r = _tn(self, *a, **k)
            return r if r.flags.writeable else r.copy()

        to_numpy._vapc_writable = True
        pd.DataFrame.to_numpy = to_numpy
    import warnings

    sys.path.insert(0, REFERENCE_SRC)
    try:
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")  # utilities.py:6 imports the deprecated pkg_resources
            import vapc  # noqa: the reference, unmodified
    finally:
        sys.path.remove(REFERENCE_SRC)
    vapc.enable_trace(False)
    vapc.enable_timeit(False)
    return vapc
