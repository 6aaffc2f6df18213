"""TEST INFRASTRUCTURE — not laspy.  Just enough of laspy's READING interface for the reference's integration tests
(tests/integration_tests/test_main.py, test_io.py) to inspect the files the drop-in package wrote, served by the repository's own LAS /
LAZ reader (vapc_b200.las).  tests/test_gpu_reference_suite.py puts this directory on sys.path only AFTER vapc_b200 has been imported, so
the package under test never sees it (its DataHandler keeps its native path)."""
from types import SimpleNamespace

from vapc_b200 import las as _las

from . import errors  # noqa: F401


class _Points:
    def __init__(self, cols, n):
        self._cols, self._n = cols, n

    def __len__(self):
        return self._n

    def __getitem__(self, name):
        return self._cols[name]  # KeyError for a dimension the file does not hold (laspy: ValueError; both fail the reference's asserts)


def read(path):
    header, cols = _las.read_las(str(path))
    fmt = SimpleNamespace(id=int(header.point_format), dimension_names=list(cols))
    hdr = SimpleNamespace(major_version=int(header.version[0]), minor_version=int(header.version[1]), point_format=fmt,
                          point_count=int(header.point_count), scales=header.scales, offsets=header.offsets)
    return SimpleNamespace(points=_Points(cols, int(header.point_count)), header=hdr, point_format=fmt, x=cols["x"], y=cols["y"], z=cols["z"])
