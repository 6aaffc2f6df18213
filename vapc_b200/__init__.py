"""vapc_b200 — B200-native (sm_100a CUDA) implementation of the voxelisation and per-voxel
statistics path of 3dgeo-heidelberg/vapc, behind the reference's Python API.

    from vapc_b200 import Vapc, BiTemporalVapc          # drop-in for vapc.Vapc / vapc.BiTemporalVapc
    from vapc_b200.engine import VoxelCloud             # device-level API (torch CUDA tensors)

All computation goes through libvapc_b200.so (C ABI in include/vapc_b200.h).  There is no CPU
fallback: importing the package without the built library raises ImportError, and using it
without a CUDA device raises RuntimeError.
"""
from . import _lib  # noqa: F401  (fails loudly when the CUDA library is missing)
from .engine import VoxelCloud, VoxelMask, FEATURE_NAMES  # noqa: F401
from .utilities import enable_trace, enable_timeit, trace, timeit, DECORATOR_CONFIG  # noqa: F401
from .vapc import Vapc  # noqa: F401
from .bi_vapc import BiTemporalVapc  # noqa: F401
from .datahandler import DataHandler  # noqa: F401
from . import vapc_tools  # noqa: F401
from .vapc_tools import use_tool  # noqa: F401
from .main import do_vapc_on_files, main, get_version  # noqa: F401

__all__ = ["Vapc", "BiTemporalVapc", "DataHandler", "use_tool", "do_vapc_on_files", "main", "get_version", "VoxelCloud", "VoxelMask",
           "enable_trace", "enable_timeit", "trace", "timeit"]
