"""Import harness for the UNMODIFIED reference (3dgeo-heidelberg/vapc) in THIS container.

TEST INFRASTRUCTURE ONLY.  Used by `oracle/make_golden.py` to generate the committed
fixtures under `tests/golden/`.  It needs `/root/reference` (absent on the GPU box), so
nothing in `tests/ -m gpu`, `bench.py` or `__graft_entry__.smoke()` may import it.

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

REFERENCE_SRC = os.environ.get("VAPC_REFERENCE_SRC", "/root/reference/src")


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
    sys.path.insert(0, REFERENCE_SRC)
    try:
        import vapc  # noqa: the reference, unmodified
    finally:
        sys.path.remove(REFERENCE_SRC)
    vapc.enable_trace(False)
    vapc.enable_timeit(False)
    return vapc
