"""trace / timeit decorators with the reference's switches (src/vapc/utilities.py:51-111).

Same names and printing behaviour, so notebooks calling `vapc.enable_trace(False)` keep working.
`timeit` synchronises the CUDA device before reading the clock on both sides — kernels are
asynchronous, and without it the printed time would only be the launch time.  Unlike the
reference both switches default to off-by-environment: set VAPC_TRACE=1 / VAPC_TIMEIT=1 or call
enable_trace() / enable_timeit() to get the reference's chatty default.
"""
import os
import time
from functools import wraps

DECORATOR_CONFIG = {"trace": os.environ.get("VAPC_TRACE", "0") == "1", "timeit": os.environ.get("VAPC_TIMEIT", "0") == "1"}


def enable_trace(enable=True):
    DECORATOR_CONFIG["trace"] = enable


def enable_timeit(enable=True):
    DECORATOR_CONFIG["timeit"] = enable


def _sync():
    try:
        import torch

        if torch.cuda.is_available():
            torch.cuda.synchronize()
    except Exception:  # pragma: no cover
        pass


def trace(func):
    @wraps(func)
    def wrapper(*args, **kwargs):
        if DECORATOR_CONFIG["trace"]:
            print(f"Calling {func.__name__}")
        return func(*args, **kwargs)

    return wrapper


def timeit(func):
    @wraps(func)
    def wrapper(*args, **kwargs):
        if not DECORATOR_CONFIG["timeit"]:
            return func(*args, **kwargs)
        _sync()
        t0 = time.time()
        result = func(*args, **kwargs)
        _sync()
        print(f"Function '{func.__name__}' executed in {time.time() - t0:.4f} seconds")
        return result

    return wrapper


def compute_mode_continuous(data):
    """Mode of a continuous sample as the centre of the fullest bin of a Freedman-Diaconis histogram (the reference's
    utilities.py:114-159, which `mode` falls back to for voxels whose values np.bincount rejects, vapc.py:407-412).  Raises like the
    reference when the bin width degenerates (a single value, zero inter-quartile range, NaN): int(ceil(nan)) / int(ceil(inf))."""
    import numpy as np

    data = np.asarray(data, dtype=np.float64)
    n = len(data)
    iqr = np.subtract(*np.percentile(data, [75, 25]))
    with np.errstate(all="ignore"):
        bin_width = 2 * iqr / n ** (1 / 3)
        bins = int(np.ceil((np.max(data) - np.min(data)) / bin_width))
    counts, bin_edges = np.histogram(data, bins=bins, density=True)
    k = int(np.argmax(counts))
    return (bin_edges[k] + bin_edges[k + 1]) / 2


def get_version():
    """The package version (the reference keeps this function here, utilities.py:11-12; `from vapc.utilities import *` exports it)."""
    from .main import __version__

    return __version__


def optimal_bins(data, method="fd"):
    """Number of histogram bins by the sqrt / Sturges / Rice / Freedman-Diaconis / Scott rule (the reference's helper, utilities.py:132-
    159; host arithmetic on one small sample, used by nothing on the device path)."""
    import numpy as np

    n = len(data)
    if method == "sqrt":
        return int(np.sqrt(n))
    if method == "sturges":
        return int(np.ceil(np.log2(n) + 1))
    if method == "rice":
        return int(np.ceil(2 * n ** (1 / 3)))
    if method == "fd":
        iqr = np.subtract(*np.percentile(data, [75, 25]))
        return int(np.ceil((np.max(data) - np.min(data)) / (2 * iqr / n ** (1 / 3))))
    if method == "scott":
        return int(np.ceil((np.max(data) - np.min(data)) / (3.5 * np.std(data) / n ** (1 / 3))))
    raise ValueError(f"Unknown method: {method}")
