"""Minimal `plyfile` for this image (the real package is not installed): the PlyData / PlyElement surface the reference's
utils/pc_utils.py uses (:130,165,218,281) -- PlyElement.describe(structured_array, name), PlyData(elements, text=False).write(path),
PlyData.read(path)[name].data -- implemented from the PLY format definition (header of `element` / `property` lines, then the
records in ascii or binary little/big endian).  Scalar properties and one-dimensional fixed-length sub-array fields (written as
`property list uchar <type>`, what plyfile does for e.g. vertex_indices) are supported; that covers every file the reference
writes and the .ply inputs it reads."""
import numpy as np

_TO_PLY = {"i1": "char", "u1": "uchar", "i2": "short", "u2": "ushort", "i4": "int", "u4": "uint", "f4": "float", "f8": "double"}
_FROM_PLY = {v: k for k, v in _TO_PLY.items()}
_FROM_PLY.update({"int8": "i1", "uint8": "u1", "int16": "i2", "uint16": "u2", "int32": "i4", "uint32": "u4", "float32": "f4",
                  "float64": "f8"})


class PlyElement:
    def __init__(self, name, data):
        self.name, self.data = name, data

    @staticmethod
    def describe(data, name, **_ignored):
        if not isinstance(data, np.ndarray) or data.dtype.names is None or data.ndim != 1:
            raise ValueError("PlyElement.describe: a one-dimensional structured array is required")
        return PlyElement(name, data)

    @property
    def count(self):
        return len(self.data)

    def __getitem__(self, key):
        return self.data[key]

    def _header(self):
        lines = ["element %s %d" % (self.name, len(self.data))]
        for field in self.data.dtype.names:
            dt = self.data.dtype[field]
            if dt.shape:                       # fixed-length sub-array -> list property
                lines.append("property list uchar %s %s" % (_TO_PLY[dt.base.str[1:]], field))
            else:
                lines.append("property %s %s" % (_TO_PLY[dt.str[1:]], field))
        return lines


class PlyData:
    def __init__(self, elements=(), text=False, byte_order="<", comments=(), obj_info=()):
        self.elements, self.text, self.byte_order = list(elements), bool(text), byte_order
        self.comments = list(comments)

    def __getitem__(self, name):
        for e in self.elements:
            if e.name == name:
                return e
        raise KeyError(name)

    def __contains__(self, name):
        return any(e.name == name for e in self.elements)

    # ---- writing ---------------------------------------------------------------------------------------------------
    def write(self, stream):
        own = isinstance(stream, (str, bytes))
        f = open(stream, "wb") if own else stream
        try:
            fmt = "ascii" if self.text else ("binary_little_endian" if self.byte_order == "<" else "binary_big_endian")
            header = ["ply", "format %s 1.0" % fmt] + ["comment " + c for c in self.comments]
            for e in self.elements:
                header += e._header()
            header.append("end_header")
            f.write(("\n".join(header) + "\n").encode("ascii"))
            for e in self.elements:
                names = e.data.dtype.names
                if self.text:
                    for rec in e.data:
                        parts = []
                        for n in names:
                            v = rec[n]
                            if np.ndim(v):
                                parts.append(str(len(v)))
                                parts += [repr(x.item()) for x in v]
                            else:
                                parts.append(repr(v.item()))
                        f.write((" ".join(parts) + "\n").encode("ascii"))
                else:
                    desc = []
                    for n in names:
                        dt = e.data.dtype[n]
                        if dt.shape:
                            desc.append((n + "__len", "u1"))
                            desc.append((n, self.byte_order + dt.base.str[1:], dt.shape))
                        else:
                            desc.append((n, self.byte_order + dt.str[1:]))
                    out = np.empty(len(e.data), dtype=desc)
                    for n in names:
                        out[n] = e.data[n]
                        if e.data.dtype[n].shape:
                            out[n + "__len"] = e.data.dtype[n].shape[0]
                    f.write(out.tobytes())
        finally:
            if own:
                f.close()

    # ---- reading ---------------------------------------------------------------------------------------------------
    @staticmethod
    def read(stream):
        own = isinstance(stream, (str, bytes))
        f = open(stream, "rb") if own else stream
        try:
            if f.readline().strip() != b"ply":
                raise ValueError("not a PLY file")
            fmt, elements, comments = None, [], []
            while True:
                line = f.readline()
                if not line:
                    raise ValueError("truncated PLY header")
                tok = line.decode("ascii", "replace").split()
                if not tok:
                    continue
                if tok[0] == "format":
                    fmt = tok[1]
                elif tok[0] == "comment":
                    comments.append(" ".join(tok[1:]))
                elif tok[0] == "element":
                    elements.append([tok[1], int(tok[2]), []])
                elif tok[0] == "property":
                    if tok[1] == "list":
                        elements[-1][2].append((tok[4], _FROM_PLY[tok[3]], _FROM_PLY[tok[2]]))
                    else:
                        elements[-1][2].append((tok[2], _FROM_PLY[tok[1]], None))
                elif tok[0] == "end_header":
                    break
            order = ">" if fmt == "binary_big_endian" else "<"
            out = []
            for name, count, props in elements:
                has_list = any(p[2] is not None for p in props)
                if fmt == "ascii":
                    rows = [f.readline().split() for _ in range(count)]
                    if not has_list:
                        arr = np.empty(count, dtype=[(n, t) for n, t, _ in props])
                        for i, n in enumerate(arr.dtype.names):
                            arr[n] = [r[i] for r in rows]
                    else:
                        arr = PlyData._lists_from_rows(rows, props)
                elif not has_list:
                    arr = np.frombuffer(f.read(count * np.dtype([(n, order + t) for n, t, _ in props]).itemsize),
                                        dtype=[(n, order + t) for n, t, _ in props]).astype([(n, t) for n, t, _ in props])
                else:
                    arr = PlyData._lists_binary(f, count, props, order)
                out.append(PlyElement(name, arr))
            return PlyData(out, text=fmt == "ascii", byte_order=order, comments=comments)
        finally:
            if own:
                f.close()

    @staticmethod
    def _lists_from_rows(rows, props):
        cols = {n: [] for n, _, _ in props}
        for r in rows:
            i = 0
            for n, t, lt in props:
                if lt is None:
                    cols[n].append(np.dtype(t).type(r[i])); i += 1
                else:
                    k = int(r[i]); cols[n].append(np.array(r[i + 1:i + 1 + k], dtype=t)); i += 1 + k
        return PlyData._assemble(cols, props, len(rows))

    @staticmethod
    def _lists_binary(f, count, props, order):
        cols = {n: [] for n, _, _ in props}
        for _ in range(count):
            for n, t, lt in props:
                if lt is None:
                    cols[n].append(np.frombuffer(f.read(np.dtype(t).itemsize), dtype=order + t)[0])
                else:
                    k = int(np.frombuffer(f.read(np.dtype(lt).itemsize), dtype=order + lt)[0])
                    cols[n].append(np.frombuffer(f.read(k * np.dtype(t).itemsize), dtype=order + t).astype(t))
        return PlyData._assemble(cols, props, count)

    @staticmethod
    def _assemble(cols, props, count):
        desc = []
        for n, t, lt in props:
            if lt is None:
                desc.append((n, t))
            else:
                lens = {len(v) for v in cols[n]}
                desc.append((n, t, (lens.pop(),)) if len(lens) == 1 else (n, object))
        arr = np.empty(count, dtype=desc)
        for n, _, _ in props:
            for i, v in enumerate(cols[n]):
                arr[n][i] = v
        return arr
