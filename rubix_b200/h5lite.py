"""Minimal read-only HDF5 reader (numpy only; the image has no h5py).

Covers what rubix's on-disk files use: superblock v0, v1 object headers, "old style" groups
(symbol-table message -> v1 B-tree -> SNOD nodes, names in a local heap), contiguous or compact
dataset layouts, little-endian fixed-point / IEEE float datatypes, and scalar string/number
attributes (rubix stores units as string attributes).  This replaces the ``h5py.File`` calls of
rubix/spectra/ssp/grid.py:323-331 and rubix/core/data.py:508-540 on the host side.

Anything else (chunked / compressed datasets, v2 headers, new-style groups) raises
``NotImplementedError`` rather than guessing.
"""

from __future__ import annotations

import struct
from typing import Dict, List, Optional, Tuple, Union

import numpy as np

_SIG = b"\x89HDF\r\n\x1a\n"
_UNDEF = 0xFFFFFFFFFFFFFFFF


class H5Dataset:
    def __init__(self, file: "H5File", name: str, shape, dtype, addr: Optional[int],
                 raw: Optional[bytes], attrs: Dict[str, object]):
        self._file = file
        self.name = name
        self.shape = tuple(shape)
        self.dtype = np.dtype(dtype)
        self._addr = addr
        self._raw = raw
        self.attrs = attrs

    def read(self) -> np.ndarray:
        n = int(np.prod(self.shape)) if self.shape else 1
        if self._raw is not None:
            arr = np.frombuffer(self._raw, dtype=self.dtype, count=n)
        elif self._addr is None or self._addr == _UNDEF:
            arr = np.zeros(n, dtype=self.dtype)
        else:
            arr = np.frombuffer(self._file._buf, dtype=self.dtype, count=n, offset=self._addr)
        if self.dtype.names == ("len", "addr", "idx"):  # variable-length strings
            f = self._file
            out = np.array([f._global_heap_object(int(r["addr"]) + f._base, int(r["idx"]))
                            .split(b"\0")[0].decode() for r in arr], dtype=object)
            return out.reshape(self.shape)
        return arr.reshape(self.shape).copy()

    def __getitem__(self, key):
        return self.read()[key]

    def __repr__(self):
        return f"<H5Dataset {self.name} shape={self.shape} dtype={self.dtype}>"


class H5Group:
    def __init__(self, file: "H5File", name: str, links: Dict[str, int], attrs: Dict[str, object]):
        self._file = file
        self.name = name
        self._links = links
        self.attrs = attrs

    def keys(self) -> List[str]:
        return list(self._links.keys())

    def __contains__(self, key: str) -> bool:
        try:
            self[key]
            return True
        except KeyError:
            return False

    def __getitem__(self, path: str) -> Union["H5Group", H5Dataset]:
        node: Union[H5Group, H5Dataset] = self
        for part in [p for p in path.split("/") if p]:
            if not isinstance(node, H5Group) or part not in node._links:
                raise KeyError(path)
            child_name = (node.name.rstrip("/") + "/" + part)
            node = node._file._load_object(node._links[part], child_name)
        return node

    def __repr__(self):
        return f"<H5Group {self.name} keys={self.keys()}>"


class H5File(H5Group):
    """``H5File(path)[...]`` mirrors the small part of ``h5py.File`` rubix uses."""

    def __init__(self, path: str):
        with open(path, "rb") as fh:
            self._buf = fh.read()
        b = self._buf
        if b[:8] != _SIG:
            raise ValueError(f"{path}: not an HDF5 file")
        if b[8] != 0:
            raise NotImplementedError(f"{path}: superblock version {b[8]} (only v0 supported)")
        if b[13] != 8 or b[14] != 8:
            raise NotImplementedError("only 8-byte offsets/lengths supported")
        self._base = struct.unpack_from("<Q", b, 24)[0]
        root_oh = struct.unpack_from("<Q", b, 56 + 8)[0]
        self._cache: Dict[int, Union[H5Group, H5Dataset]] = {}
        root = self._load_object(root_oh, "/")
        if not isinstance(root, H5Group):
            raise ValueError("root object is not a group")
        super().__init__(self, "/", root._links, root.attrs)

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        return False

    # ---- object header ---------------------------------------------------------------
    def _messages(self, addr: int) -> List[Tuple[int, bytes]]:
        b = self._buf
        if b[addr] != 1:
            raise NotImplementedError(f"object header version {b[addr]} at {addr}")
        nmsg = struct.unpack_from("<H", b, addr + 2)[0]
        hsize = struct.unpack_from("<I", b, addr + 8)[0]
        msgs: List[Tuple[int, bytes]] = []
        blocks = [(addr + 16, hsize)]
        while blocks and len(msgs) < nmsg:
            pos, size = blocks.pop(0)
            end = pos + size
            while pos + 8 <= end and len(msgs) < nmsg:
                mtype, msize, _flags = struct.unpack_from("<HHB", b, pos)
                body = b[pos + 8:pos + 8 + msize]
                pos += 8 + msize
                if mtype == 0x10:  # continuation
                    caddr, clen = struct.unpack_from("<QQ", body, 0)
                    blocks.append((caddr + self._base, clen))
                msgs.append((mtype, body))
        return msgs

    def _load_object(self, addr: int, name: str):
        addr += self._base
        if addr in self._cache:
            return self._cache[addr]
        msgs = self._messages(addr)
        attrs: Dict[str, object] = {}
        shape = dtype = None
        layout = None
        links = None
        for mtype, body in msgs:
            if mtype == 0x01:
                shape = self._dataspace(body)
            elif mtype == 0x03:
                dtype = self._datatype(body)[0]
            elif mtype == 0x08:
                layout = self._layout(body)
            elif mtype == 0x11:
                btree, heap = struct.unpack_from("<QQ", body, 0)
                links = self._group_links(btree + self._base, heap + self._base)
            elif mtype == 0x0C:
                try:
                    k, v = self._attribute(body)
                    attrs[k] = v
                except NotImplementedError:
                    pass
            elif mtype == 0x0B:
                raise NotImplementedError(f"{name}: filtered (compressed) datasets not supported")
        if links is not None:
            obj: Union[H5Group, H5Dataset] = H5Group(self, name, links, attrs)
        elif shape is not None and dtype is not None and layout is not None:
            kind, a, raw = layout
            obj = H5Dataset(self, name, shape, dtype, a, raw, attrs)
        else:
            raise NotImplementedError(f"{name}: unsupported object (no symbol table / dataset)")
        self._cache[addr] = obj
        return obj

    # ---- messages ----------------------------------------------------------------------
    @staticmethod
    def _dataspace(body: bytes):
        ver, rank, flags = body[0], body[1], body[2]
        if ver == 1:
            off = 8
        elif ver == 2:
            off = 4
        else:
            raise NotImplementedError(f"dataspace version {ver}")
        return struct.unpack_from("<" + "Q" * rank, body, off) if rank else ()

    @staticmethod
    def _datatype(body: bytes):
        cv = body[0]
        cls, ver = cv & 0x0F, cv >> 4
        bits0 = body[1]
        size = struct.unpack_from("<I", body, 4)[0]
        if bits0 & 1 and cls in (0, 1):
            raise NotImplementedError("big-endian data")
        if cls == 0:  # fixed point
            signed = bool(bits0 & 0x08)
            return np.dtype(("<i" if signed else "<u") + str(size)), size
        if cls == 1:  # float
            return np.dtype("<f" + str(size)), size
        if cls == 3:  # fixed-length string
            return np.dtype("S" + str(size)), size
        if cls == 9:  # variable length (strings): 16-byte global-heap references
            return np.dtype([("len", "<u4"), ("addr", "<u8"), ("idx", "<u4")]), 16
        raise NotImplementedError(f"datatype class {cls}")

    def _layout(self, body: bytes):
        ver = body[0]
        if ver == 3:
            cls = body[1]
            if cls == 1:
                a, _size = struct.unpack_from("<QQ", body, 2)
                return ("contiguous", None if a == _UNDEF else a + self._base, None)
            if cls == 0:
                n = struct.unpack_from("<H", body, 2)[0]
                return ("compact", None, bytes(body[4:4 + n]))
            raise NotImplementedError("chunked dataset layout")
        if ver in (1, 2):
            rank, cls = body[1], body[2]
            if cls == 1:
                a = struct.unpack_from("<Q", body, 8)[0]
                return ("contiguous", None if a == _UNDEF else a + self._base, None)
            raise NotImplementedError("layout v1/v2 non-contiguous")
        raise NotImplementedError(f"layout version {ver}")

    def _attribute(self, body: bytes):
        ver = body[0]
        if ver not in (1, 2, 3):
            raise NotImplementedError
        nsz, tsz, ssz = struct.unpack_from("<HHH", body, 2)
        pos = 8
        if ver == 3:
            pos += 1
        pad = (lambda n: (n + 7) & ~7) if ver == 1 else (lambda n: n)
        name = body[pos:pos + nsz].split(b"\0")[0].decode()
        pos += pad(nsz)
        tbody = body[pos:pos + tsz]
        pos += pad(tsz)
        sbody = body[pos:pos + ssz]
        pos += pad(ssz)
        cls = tbody[0] & 0x0F
        if cls == 9:  # variable-length (string) -> global heap reference
            _n, gaddr, gidx = struct.unpack_from("<IQI", body, pos)
            return name, self._global_heap_object(gaddr + self._base, gidx).split(b"\0")[0].decode()
        dtype, size = self._datatype(tbody)
        shape = self._dataspace(sbody) if ssz >= 4 else ()
        n = int(np.prod(shape)) if shape else 1
        arr = np.frombuffer(body, dtype=dtype, count=n, offset=pos)
        if dtype.kind == "S":
            val = arr[0].split(b"\0")[0].decode()
            return name, val
        return name, (arr.reshape(shape).copy() if shape else arr[0])

    def _global_heap_object(self, addr: int, index: int) -> bytes:
        b = self._buf
        if b[addr:addr + 4] != b"GCOL":
            raise NotImplementedError("bad global heap")
        size = struct.unpack_from("<Q", b, addr + 8)[0]
        pos, end = addr + 16, addr + size
        while pos + 16 <= end:
            idx, _rc, _r, osz = struct.unpack_from("<HHIQ", b, pos)
            if idx == index:
                return bytes(b[pos + 16:pos + 16 + osz])
            if idx == 0:
                break
            pos += 16 + ((osz + 7) & ~7)
        raise KeyError(index)

    # ---- groups ------------------------------------------------------------------------
    def _group_links(self, btree: int, heap: int) -> Dict[str, int]:
        b = self._buf
        if b[heap:heap + 4] != b"HEAP":
            raise NotImplementedError("bad local heap")
        data_addr = struct.unpack_from("<Q", b, heap + 24)[0] + self._base
        links: Dict[str, int] = {}

        def name_at(off):
            s = data_addr + off
            e = b.index(b"\0", s)
            return b[s:e].decode()

        def walk(node):
            if b[node:node + 4] == b"TREE":
                ntype, level, used = struct.unpack_from("<BBH", b, node + 4)
                if ntype != 0:
                    raise NotImplementedError("non-group B-tree")
                pos = node + 24  # after sig(4)+type/level/used(4)+siblings(16)
                for k in range(used):
                    child = struct.unpack_from("<Q", b, pos + 8)[0] + self._base
                    pos += 16
                    walk(child)
            elif b[node:node + 4] == b"SNOD":
                nsym = struct.unpack_from("<H", b, node + 6)[0]
                pos = node + 8
                for _ in range(nsym):
                    noff, oh = struct.unpack_from("<QQ", b, pos)
                    links[name_at(noff)] = oh
                    pos += 40
            else:
                raise NotImplementedError("unexpected group node")

        walk(btree)
        return links


def visit(group: H5Group, prefix: str = ""):
    """Yield ``(path, object)`` for every dataset below ``group``."""
    for k in group.keys():
        obj = group[k]
        p = prefix + "/" + k
        if isinstance(obj, H5Group):
            yield from visit(obj, p)
        else:
            yield p, obj
