"""Minimal `plyfile` stand-in: binary little-endian PLY read / write of fixed-width scalar properties —
what the reference's checkpoint I/O needs (scene/gaussian_model.py:761-855 `save_ply*`: one 'vertex'
element built with `PlyElement.describe(structured_array, 'vertex')` and written with
`PlyData([el]).write(path)`; :934-1027 `load_ply`: `PlyData.read(path)`, `plydata.elements[0]["x"]`,
`plydata.elements[0].properties` -> `.name`; scene/dataset_readers.py:234-260 `fetchPly` / `storePly`:
`plydata['vertex']`, `vertices['x']`).  Files are interchangeable with the real package for this subset
(same header grammar, same packed record layout).  ASCII files and list properties are read too, since
point clouds written by other tools use them; they are never written.

Only used when the real `plyfile` is absent (compat/ is appended to sys.path, see compat/README.md)."""
from __future__ import annotations

import numpy as np

_PLY2NP = {"char": "i1", "int8": "i1", "uchar": "u1", "uint8": "u1", "short": "i2", "int16": "i2",
           "ushort": "u2", "uint16": "u2", "int": "i4", "int32": "i4", "uint": "u4", "uint32": "u4",
           "float": "f4", "float32": "f4", "double": "f8", "float64": "f8"}
_NP2PLY = {"i1": "char", "u1": "uchar", "i2": "short", "u2": "ushort", "i4": "int", "u4": "uint",
           "f4": "float", "f8": "double"}


class PlyProperty:
    def __init__(self, name, val_dtype):
        self.name, self.val_dtype = name, val_dtype

    def __repr__(self):
        return f"PlyProperty({self.name!r}, {self.val_dtype!r})"


class PlyListProperty(PlyProperty):
    def __init__(self, name, len_dtype, val_dtype):
        super().__init__(name, val_dtype)
        self.len_dtype = len_dtype


class PlyElement:
    def __init__(self, name, properties, data):
        self.name, self.properties, self.data = name, tuple(properties), data

    @staticmethod
    def describe(data, name, **_unused):
        data = np.asarray(data)
        if data.dtype.names is None:
            raise ValueError("PlyElement.describe needs a structured array")
        props = []
        for n in data.dtype.names:
            dt = data.dtype[n]
            key = dt.str.lstrip("<>|=")
            if dt.shape != () or key not in _NP2PLY:
                raise ValueError(f"property {n!r}: dtype {dt} is outside the supported scalar subset")
            props.append(PlyProperty(n, key))
        return PlyElement(name, props, data)

    @property
    def count(self):
        return len(self.data)

    def __len__(self):
        return len(self.data)

    def __getitem__(self, key):
        return self.data[key]


class PlyData:
    def __init__(self, elements=(), text=False, byte_order="<", comments=()):
        self.elements = list(elements)
        self.text, self.comments = text, list(comments)

    def __getitem__(self, name):
        for e in self.elements:
            if e.name == name:
                return e
        raise KeyError(name)

    def __contains__(self, name):
        return any(e.name == name for e in self.elements)

    # ------------------------------------------------------------------ write
    def write(self, stream):
        own = isinstance(stream, (str, bytes)) or hasattr(stream, "__fspath__")
        f = open(stream, "wb") if own else stream
        try:
            head = ["ply", "format binary_little_endian 1.0"] + [f"comment {c}" for c in self.comments]
            for e in self.elements:
                head.append(f"element {e.name} {len(e.data)}")
                head += [f"property {_NP2PLY[p.val_dtype]} {p.name}" for p in e.properties]
            head.append("end_header")
            f.write(("\n".join(head) + "\n").encode("ascii"))
            for e in self.elements:
                packed = np.dtype([(p.name, "<" + p.val_dtype) for p in e.properties])      # no padding
                out = np.empty(len(e.data), dtype=packed)
                for p in e.properties:
                    out[p.name] = e.data[p.name]
                f.write(out.tobytes())
        finally:
            if own:
                f.close()

    # ------------------------------------------------------------------ read
    @staticmethod
    def read(stream):
        own = isinstance(stream, (str, bytes)) or hasattr(stream, "__fspath__")
        f = open(stream, "rb") if own else stream
        try:
            if f.readline().strip() != b"ply":
                raise ValueError("not a PLY file")
            fmt, comments, elems = None, [], []
            while True:
                line = f.readline()
                if not line:
                    raise ValueError("PLY header without end_header")
                tok = line.decode("ascii").split()
                if not tok:
                    continue
                if tok[0] == "format":
                    fmt = tok[1]
                elif tok[0] in ("comment", "obj_info"):
                    comments.append(" ".join(tok[1:]))
                elif tok[0] == "element":
                    elems.append((tok[1], int(tok[2]), []))
                elif tok[0] == "property":
                    if tok[1] == "list":
                        elems[-1][2].append(PlyListProperty(tok[4], _PLY2NP[tok[2]], _PLY2NP[tok[3]]))
                    else:
                        elems[-1][2].append(PlyProperty(tok[2], _PLY2NP[tok[1]]))
                elif tok[0] == "end_header":
                    break
            if fmt not in ("binary_little_endian", "binary_big_endian", "ascii"):
                raise ValueError(f"unsupported PLY format {fmt!r}")
            bo = ">" if fmt == "binary_big_endian" else "<"
            out = []
            for name, count, props in elems:
                has_list = any(isinstance(p, PlyListProperty) for p in props)
                if fmt != "ascii" and not has_list:
                    dt = np.dtype([(p.name, bo + p.val_dtype) for p in props])
                    raw = f.read(dt.itemsize * count)
                    if len(raw) != dt.itemsize * count:
                        raise ValueError(f"PLY element {name!r} is truncated")
                    data = np.frombuffer(raw, dtype=dt, count=count).astype(dt.newbyteorder("="), copy=True)
                else:
                    data = PlyData._read_slow(f, fmt, bo, count, props)
                out.append(PlyElement(name, props, data))
            return PlyData(out, text=(fmt == "ascii"), comments=comments)
        finally:
            if own:
                f.close()

    @staticmethod
    def _read_slow(f, fmt, bo, count, props):
        dt = np.dtype([(p.name, object if isinstance(p, PlyListProperty) else p.val_dtype) for p in props])
        data = np.empty(count, dtype=dt)
        for i in range(count):
            if fmt == "ascii":
                tok = f.readline().split()
                pos = 0
                for p in props:
                    if isinstance(p, PlyListProperty):
                        n = int(tok[pos]); pos += 1
                        data[p.name][i] = np.array(tok[pos:pos + n], dtype=np.float64).astype(p.val_dtype)
                        pos += n
                    else:
                        data[p.name][i] = np.array(tok[pos], dtype=np.float64).astype(p.val_dtype)
                        pos += 1
            else:
                for p in props:
                    if isinstance(p, PlyListProperty):
                        ld = np.dtype(bo + p.len_dtype)
                        n = int(np.frombuffer(f.read(ld.itemsize), dtype=ld)[0])
                        vd = np.dtype(bo + p.val_dtype)
                        data[p.name][i] = np.frombuffer(f.read(vd.itemsize * n), dtype=vd).copy()
                    else:
                        vd = np.dtype(bo + p.val_dtype)
                        data[p.name][i] = np.frombuffer(f.read(vd.itemsize), dtype=vd)[0]
        return data
