"""Writer / reader of the container files oracle/h5shim reads and writes -- TEST / BASELINE
INFRASTRUCTURE (the reference built with oracle/h5shim sees these files through its HDF5
calls; real HDF5 is not in this image).  The API follows the small part of h5py that
openmc's exporters use (openmc/data/neutron.py:354-431, openmc/mgxs_library.py:2504+):

    with omh5.File(path, "w") as f:
        g = f.create_group("U235"); g.attrs["Z"] = 92
        g.create_dataset("energy/294K", data=array)
    f = omh5.File(path, "r"); f["U235/energy/294K"][()]; f["U235"].attrs["Z"]

Layout of the file: oracle/h5shim/h5shim.cpp (header comment).
"""
from __future__ import annotations

import struct

import numpy as np

MAGIC = b"OMH5\x01\x00\x00\x00"
_CLS = {"i": 0, "u": 1, "f": 2, "S": 3, "V": 4, "b": 0}


def _as_array(value):
    if isinstance(value, str):
        value = value.encode()
    if isinstance(value, bytes):
        return np.array(value, dtype="S%d" % max(len(value), 1))
    a = np.asarray(value)
    if a.dtype.kind == "U":
        a = np.char.encode(a)
    if a.dtype.kind == "b":
        a = a.astype(np.int8)
    if a.dtype.kind == "O":
        raise TypeError("object arrays cannot be stored")
    return np.ascontiguousarray(a)


def _type_bytes(dt: np.dtype) -> bytes:
    if dt.kind == "V" and dt.names:
        out = struct.pack("<BQI", 4, dt.itemsize, len(dt.names))
        for n in dt.names:
            sub, off = dt.fields[n][:2]
            nb = n.encode()
            out += struct.pack("<I", len(nb)) + nb + struct.pack("<Q", off) + _type_bytes(sub)
        return out
    return struct.pack("<BQ", _CLS[dt.kind], dt.itemsize)


def _data_bytes(a: np.ndarray) -> bytes:
    a = _as_array(a)
    if a.dtype.byteorder == ">":
        a = a.astype(a.dtype.newbyteorder("<"))
    raw = a.tobytes()
    head = _type_bytes(a.dtype) + struct.pack("<BI", 1 if a.ndim == 0 else 0, a.ndim)
    head += b"".join(struct.pack("<Q", d) for d in a.shape)
    return head + struct.pack("<Q", len(raw)) + raw


class Dataset:
    def __init__(self, name, data):
        self.name = name
        self.attrs = {}
        self.data = _as_array(data)

    def __getitem__(self, key):
        return self.data[key]

    @property
    def shape(self):
        return self.data.shape


class Group:
    def __init__(self, name=""):
        self.name = name
        self.attrs = {}
        self.children = {}

    def _walk(self, path, create=False):
        node = self
        parts = [p for p in path.split("/") if p]
        for p in parts[:-1]:
            if p not in node.children:
                if not create:
                    raise KeyError(path)
                node.children[p] = Group(p)
            node = node.children[p]
        return node, (parts[-1] if parts else "")

    def create_group(self, path):
        parent, leaf = self._walk(path, True)
        if leaf in parent.children:
            raise ValueError("exists: " + path)
        g = parent.children[leaf] = Group(leaf)
        return g

    def require_group(self, path):
        try:
            return self[path]
        except KeyError:
            return self.create_group(path)

    def create_dataset(self, path, data=None):
        parent, leaf = self._walk(path, True)
        if leaf in parent.children:
            raise ValueError("exists: " + path)
        d = parent.children[leaf] = Dataset(leaf, data)
        return d

    def __getitem__(self, path):
        parent, leaf = self._walk(path)
        if not leaf:
            return parent
        return parent.children[leaf]

    def __contains__(self, path):
        try:
            self[path]
            return True
        except KeyError:
            return False

    def keys(self):
        return sorted(self.children)

    # -- serialisation
    def _emit(self, out):
        nb = self.name.encode()
        out.append(struct.pack("<BI", 0, len(nb)) + nb + struct.pack("<I", len(self.attrs)))
        _emit_attrs(self.attrs, out)
        out.append(struct.pack("<I", len(self.children)))
        for k in sorted(self.children):
            self.children[k]._emit(out)


def _emit_attrs(attrs, out):
    for k in sorted(attrs):
        kb = k.encode()
        out.append(struct.pack("<I", len(kb)) + kb + _data_bytes(attrs[k]))


def _dataset_emit(self, out):
    nb = self.name.encode()
    out.append(struct.pack("<BI", 1, len(nb)) + nb + struct.pack("<I", len(self.attrs)))
    _emit_attrs(self.attrs, out)
    out.append(_data_bytes(self.data))


Dataset._emit = _dataset_emit


class _Reader:
    def __init__(self, buf):
        self.b = memoryview(buf)
        self.p = 0

    def take(self, fmt):
        v = struct.unpack_from("<" + fmt, self.b, self.p)
        self.p += struct.calcsize("<" + fmt)
        return v if len(v) > 1 else v[0]

    def string(self):
        n = self.take("I")
        s = bytes(self.b[self.p:self.p + n]).decode()
        self.p += n
        return s

    def dtype(self):
        cls, size = self.take("BQ")
        if cls == 4:
            n = self.take("I")
            names, formats, offsets = [], [], []
            for _ in range(n):
                names.append(self.string())
                offsets.append(self.take("Q"))
                formats.append(self.dtype())
            return np.dtype({"names": names, "formats": formats, "offsets": offsets, "itemsize": size})
        return np.dtype({0: "<i%d", 1: "<u%d", 2: "<f%d", 3: "S%d"}[cls] % size)

    def data(self):
        dt = self.dtype()
        scalar = self.take("B")
        rank = self.take("I")
        dims = tuple(self.take("Q") for _ in range(rank))
        nb = self.take("Q")
        a = np.frombuffer(self.b, dtype=dt, count=nb // dt.itemsize if dt.itemsize else 0, offset=self.p)
        self.p += nb
        return a.reshape(dims) if not scalar or rank else a.reshape(())

    def node(self):
        kind = self.take("B")
        name = self.string()
        n_attr = self.take("I")
        attrs = {}
        for _ in range(n_attr):
            k = self.string()
            attrs[k] = self.data()
        if kind == 1:
            d = Dataset(name, self.data())
            d.attrs = attrs
            return d
        g = Group(name)
        g.attrs = attrs
        for _ in range(self.take("I")):
            c = self.node()
            g.children[c.name] = c
        return g


class File(Group):
    def __init__(self, path, mode="r"):
        super().__init__("")
        self.path, self.mode = path, mode
        if mode == "r":
            with open(path, "rb") as f:
                buf = f.read()
            if buf[:8] != MAGIC:
                raise OSError("not an OMH5 container: " + path)
            root = _Reader(buf[8:]).node()
            self.attrs, self.children = root.attrs, root.children

    def close(self):
        if self.mode == "w":
            out = [MAGIC]
            self._emit(out)
            with open(self.path, "wb") as f:
                for piece in out:
                    f.write(piece)
            self.mode = "closed"

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()
