"""TEST INFRASTRUCTURE — not plyfile.  Reads the binary little-endian PLY files the DataHandler writes (vertex element of scalar
properties + triangle faces) into the few attributes the reference's integration tests look at: elements[i].properties[j].name and
elements[i].data (a numpy structured array)."""
from types import SimpleNamespace

import numpy as np

_TYPES = {"char": "i1", "uchar": "u1", "short": "<i2", "ushort": "<u2", "int": "<i4", "uint": "<u4", "float": "<f4", "double": "<f8",
          "int8": "i1", "uint8": "u1", "int16": "<i2", "uint16": "<u2", "int32": "<i4", "uint32": "<u4", "float32": "<f4", "float64": "<f8"}


class PlyData:
    def __init__(self, elements):
        self.elements = elements

    @staticmethod
    def read(path):
        buf = open(str(path), "rb").read()
        end = buf.index(b"end_header\n") + len(b"end_header\n")
        lines = buf[:end].decode("ascii").splitlines()
        assert lines[0] == "ply" and lines[1].startswith("format binary_little_endian"), lines[:2]
        specs = []
        for ln in lines[2:]:
            w = ln.split()
            if w[0] == "element":
                specs.append((w[1], int(w[2]), []))
            elif w[0] == "property":
                specs[-1][2].append(w[1:])
        pos, elements = end, []
        for name, count, props in specs:
            fields = []
            for p in props:
                if p[0] == "list":  # property list <count type> <item type> <name>: the writer's faces are triangles
                    fields += [(p[3] + "_n", _TYPES[p[1]]), (p[3], _TYPES[p[2]], (3,))]
                else:
                    fields.append((p[1], _TYPES[p[0]]))
            dt = np.dtype(fields)
            data = np.frombuffer(buf, dtype=dt, count=count, offset=pos)
            pos += count * dt.itemsize
            elements.append(SimpleNamespace(name=name, data=data, properties=[SimpleNamespace(name=p[-1]) for p in props]))
        assert pos == len(buf), (pos, len(buf))
        return PlyData(elements)
