"""
Minimal HDF5 reader / writer for the S^3 output files (no h5py in this image, and the I/O layer should not need it).

The reference writes its results with h5py defaults (sparseSpatialSampling/data.py:346-449): superblock version 0,
old-style groups (symbol table = v1 B-tree + local heap + symbol nodes), version-1 object headers, contiguous datasets of
little-endian integers / IEEE floats. That subset of the HDF5 file format (HDF5 File Format Specification, version 1.1
structures) is what this module speaks -- enough to read the reference's own files (pinned by its golden
``tests/s_cube_test_dataset.h5``) and to write files that h5py / ParaView / the reference's ``Dataloader`` open.

Writing is append-only: the raw data of a dataset goes to the end of the file when ``write`` is called (nothing is kept
in memory), the metadata (object headers, heaps, B-trees) is written when the file is closed and the superblock is
pointed at it. Re-opening in mode ``"a"`` parses the existing tree, appends new datasets and writes a fresh metadata
block; the data already in the file is never moved or re-read, so a batch-wise export costs I/O linear in its size.
"""
import os
import struct
from typing import Dict, List, Tuple

import numpy as np

SIGNATURE = b"\x89HDF\r\n\x1a\n"
UNDEF = 0xFFFFFFFFFFFFFFFF
LEAF_K, INTERNAL_K = 4, 16                 # h5py / libhdf5 defaults: 8 symbols per node, 32 children per B-tree node
SNOD_SIZE = 8 + 2 * LEAF_K * 40
TREE_SIZE = 24 + (2 * INTERNAL_K + 1) * 8 + 2 * INTERNAL_K * 8

MSG_NIL, MSG_DATASPACE, MSG_DATATYPE, MSG_FILL_OLD, MSG_FILL, MSG_LAYOUT = 0x0, 0x1, 0x3, 0x4, 0x5, 0x8
MSG_CONTINUATION, MSG_SYMBOL_TABLE, MSG_MTIME = 0x10, 0x11, 0x12


class H5Error(RuntimeError):
    pass


# ------------------------------------------------------------------------------------------------------- datatypes
def _encode_dtype(dt: np.dtype) -> bytes:
    """Datatype message (version 1) of a little-endian fixed-point or IEEE floating-point type."""
    dt = np.dtype(dt)
    if dt.kind in "iu" or dt.kind == "b":
        size = dt.itemsize
        bits0 = 0x08 if dt.kind == "i" else 0x00              # bit 3: signed (two's complement)
        return struct.pack("<BBBBI", 0x10 | 0, bits0, 0, 0, size) + struct.pack("<HH", 0, size * 8)
    if dt.kind == "f" and dt.itemsize in (4, 8):
        size = dt.itemsize
        # bits 0: byte order LE, 1-3: padding 0, 4-5: mantissa normalisation 2 (implied msb), 8-15: sign bit position
        sign = size * 8 - 1
        bits0, bits1 = 0x20, sign
        if size == 4:
            props = struct.pack("<HHBBBBI", 0, 32, 23, 8, 0, 23, 127)
        else:
            props = struct.pack("<HHBBBBI", 0, 64, 52, 11, 0, 52, 1023)
        return struct.pack("<BBBBI", 0x10 | 1, bits0, bits1, 0, size) + props
    raise H5Error(f"h5lite cannot store dtype {dt}")


def _decode_dtype(msg: bytes) -> np.dtype:
    cls, bits0, bits1, _, size = struct.unpack_from("<BBBBI", msg, 0)
    klass = cls & 0x0F
    order = ">" if (bits0 & 1) else "<"
    if klass == 0:
        signed = bool(bits0 & 0x08)
        return np.dtype(f"{order}{'i' if signed else 'u'}{size}")
    if klass == 1:
        if size not in (2, 4, 8):
            raise H5Error(f"unsupported float size {size}")
        return np.dtype(f"{order}f{size}")
    raise H5Error(f"unsupported HDF5 datatype class {klass}")


# ------------------------------------------------------------------------------------------------------- reading
class _Dataset:
    __slots__ = ("shape", "dtype", "address", "nbytes", "compact")

    def __init__(self, shape, dtype, address, nbytes, compact=None):
        self.shape, self.dtype, self.address, self.nbytes, self.compact = shape, dtype, address, nbytes, compact


class _Reader:
    def __init__(self, f):
        self.f = f
        f.seek(0)
        head = f.read(96)
        if len(head) < 96 or head[:8] != SIGNATURE:
            raise H5Error("not an HDF5 file")
        if head[8] != 0:
            raise H5Error(f"superblock version {head[8]} is not supported (h5lite reads version 0 files)")
        if head[13] != 8 or head[14] != 8:
            raise H5Error("only 8-byte offsets and lengths are supported")
        self.base = struct.unpack_from("<Q", head, 24)[0]
        self.eof = struct.unpack_from("<Q", head, 40)[0]
        self.root_header = struct.unpack_from("<Q", head, 56 + 8)[0]

    def _read(self, addr: int, n: int) -> bytes:
        self.f.seek(self.base + addr)
        b = self.f.read(n)
        if len(b) != n:
            raise H5Error("truncated HDF5 file")
        return b

    def messages(self, addr: int) -> List[Tuple[int, bytes]]:
        """(type, data) of all messages of a version-1 object header, following continuation blocks."""
        head = self._read(addr, 16)
        version, _, n_msg, _, size = struct.unpack_from("<BBHII", head, 0)
        if version != 1:
            raise H5Error(f"object header version {version} is not supported")
        blocks = [(addr + 16, size)]
        out = []
        while blocks and len(out) < n_msg:
            start, length = blocks.pop(0)
            raw = self._read(start, length)
            pos = 0
            while pos + 8 <= length and len(out) < n_msg:
                mtype, msize, _ = struct.unpack_from("<HHB", raw, pos)
                data = raw[pos + 8:pos + 8 + msize]
                pos += 8 + msize
                out.append((mtype, data))
                if mtype == MSG_CONTINUATION:
                    off, ln = struct.unpack_from("<QQ", data, 0)
                    blocks.append((off, ln))
        return out

    def heap_names(self, heap_addr: int):
        head = self._read(heap_addr, 32)
        if head[:4] != b"HEAP":
            raise H5Error("bad local heap signature")
        size, _, data_addr = struct.unpack_from("<QQQ", head, 8)
        return self._read(data_addr, size)

    def symbols(self, btree_addr: int, heap: bytes) -> List[Tuple[str, int]]:
        """(name, object header address) of all links below a group B-tree, in name order."""
        node = self._read(btree_addr, TREE_SIZE if btree_addr + TREE_SIZE <= self.eof else self.eof - btree_addr)
        if node[:4] == b"TREE":
            node_type, level, used = struct.unpack_from("<BBH", node, 4)
            if node_type != 0:
                raise H5Error("unexpected B-tree node type")
            out = []
            for i in range(used):
                child = struct.unpack_from("<Q", node, 24 + 8 + i * 16)[0]
                out += self.symbols(child, heap)
            return out
        if node[:4] == b"SNOD":
            n = struct.unpack_from("<H", node, 6)[0]
            raw = self._read(btree_addr, 8 + n * 40)
            out = []
            for i in range(n):
                name_off, header = struct.unpack_from("<QQ", raw, 8 + i * 40)
                end = heap.index(b"\x00", name_off)
                out.append((heap[name_off:end].decode("utf-8"), header))
            return out
        raise H5Error("bad B-tree / symbol node signature")

    def tree(self, addr: int = None) -> dict:
        """Nested dict of the file: groups -> dict, datasets -> _Dataset."""
        addr = self.root_header if addr is None else addr
        msgs = self.messages(addr)
        table = [d for t, d in msgs if t == MSG_SYMBOL_TABLE]
        if table:
            btree, heap_addr = struct.unpack_from("<QQ", table[0], 0)
            heap = self.heap_names(heap_addr)
            return {name: self.tree(header) for name, header in self.symbols(btree, heap)}
        shape = dtype = layout = None
        for t, d in msgs:
            if t == MSG_DATASPACE:
                version, rank = d[0], d[1]
                off = 8 if version == 1 else 4
                shape = tuple(struct.unpack_from(f"<{rank}Q", d, off)) if rank else ()
            elif t == MSG_DATATYPE:
                dtype = _decode_dtype(d)
            elif t == MSG_LAYOUT:
                layout = d
        if shape is None or dtype is None or layout is None:
            raise H5Error("object is neither an old-style group nor a dataset h5lite understands")
        nbytes = int(np.prod(shape, dtype=np.int64)) * dtype.itemsize if shape else dtype.itemsize
        version = layout[0]
        if version == 3:
            klass = layout[1]
            if klass == 1:
                address, size = struct.unpack_from("<QQ", layout, 2)
                return _Dataset(shape, dtype, address, nbytes)
            if klass == 0:
                size = struct.unpack_from("<H", layout, 2)[0]
                return _Dataset(shape, dtype, None, nbytes, compact=bytes(layout[4:4 + size]))
            raise H5Error("chunked datasets are not supported by h5lite")
        if version in (1, 2):
            rank, klass = layout[1], layout[2]
            if klass != 1:
                raise H5Error("only contiguous datasets are supported by h5lite")
            address = struct.unpack_from("<Q", layout, 8)[0]
            return _Dataset(shape, dtype, address, nbytes)
        raise H5Error(f"data layout message version {version} is not supported")

    def read(self, ds: _Dataset) -> np.ndarray:
        if ds.compact is not None:
            raw = ds.compact
        elif ds.address == UNDEF or ds.nbytes == 0:
            raw = b"\x00" * ds.nbytes
        else:
            raw = self._read(ds.address, ds.nbytes)
        arr = np.frombuffer(raw, dtype=ds.dtype, count=ds.nbytes // ds.dtype.itemsize).reshape(ds.shape)
        return arr.astype(ds.dtype.newbyteorder("=")) if ds.dtype.byteorder == ">" else arr.copy()


# ------------------------------------------------------------------------------------------------------- writing
def _pad8(b: bytes) -> bytes:
    return b + b"\x00" * (-len(b) % 8)


def _message(mtype: int, data: bytes) -> bytes:
    data = _pad8(data)
    return struct.pack("<HHB3x", mtype, len(data), 0) + data


def _object_header(messages: List[bytes]) -> bytes:
    body = b"".join(messages)
    return struct.pack("<BBHII4x", 1, 0, len(messages), 1, len(body)) + body


class _MetaWriter:
    """Lays out the metadata block (object headers, heaps, B-trees, symbol nodes) of a tree at a given file offset."""

    def __init__(self, start: int):
        self.start = start
        self.buf = bytearray()

    def alloc(self, data: bytes) -> int:
        self.buf += b"\x00" * (-len(self.buf) % 8)
        addr = self.start + len(self.buf)
        self.buf += data
        return addr

    def dataset(self, ds: _Dataset) -> int:
        rank = len(ds.shape)
        space = struct.pack("<BBB5x", 1, rank, 0) + b"".join(struct.pack("<Q", int(n)) for n in ds.shape)
        fill = struct.pack("<BBBB", 2, 2, 2, 0)               # version 2, late allocation, write fill if set, undefined
        address = ds.address if ds.nbytes else UNDEF
        layout = struct.pack("<BBQQ", 3, 1, address, ds.nbytes)
        return self.alloc(_object_header([_message(MSG_DATASPACE, space), _message(MSG_DATATYPE, _encode_dtype(ds.dtype)),
                                          _message(MSG_FILL, fill), _message(MSG_LAYOUT, layout)]))

    def group(self, node: dict) -> Tuple[int, int, int]:
        """Writes a group; returns (object header address, B-tree address, heap address)."""
        names = sorted(node.keys(), key=lambda s: s.encode("utf-8"))              # strcmp order
        headers = {}
        for name in names:
            child = node[name]
            headers[name] = self.group(child) if isinstance(child, dict) else (self.dataset(child), None, None)
        # local heap: the empty string at offset 0, then the link names, each padded to 8 bytes
        heap = bytearray(b"\x00" * 8)
        offsets = {}
        for name in names:
            offsets[name] = len(heap)
            heap += _pad8(name.encode("utf-8") + b"\x00")
        heap_data_addr = self.alloc(bytes(heap))
        heap_addr = self.alloc(b"HEAP" + struct.pack("<B3xQQQ", 0, len(heap), 1, heap_data_addr))   # 1 = no free block
        # symbol nodes (<= 2 * LEAF_K links each), then B-tree levels (<= 2 * INTERNAL_K children each)
        level_nodes = []                                                         # (address, offset of largest name)
        for i in range(0, max(len(names), 1), 2 * LEAF_K):
            part = names[i:i + 2 * LEAF_K]
            raw = bytearray(b"SNOD" + struct.pack("<BBH", 1, 0, len(part)))
            for name in part:
                header, btree, hp = headers[name]
                if btree is not None:                                            # cached group info in the scratch pad
                    raw += struct.pack("<QQII", offsets[name], header, 1, 0) + struct.pack("<QQ", btree, hp)
                else:
                    raw += struct.pack("<QQII16x", offsets[name], header, 0, 0)
            raw += b"\x00" * (SNOD_SIZE - len(raw))
            level_nodes.append((self.alloc(bytes(raw)), offsets[part[-1]] if part else 0))
        level = 0
        while True:
            parents = []
            for i in range(0, len(level_nodes), 2 * INTERNAL_K):
                part = level_nodes[i:i + 2 * INTERNAL_K]
                parents.append((part, level))
            nodes = []
            addrs = []
            for part, lvl in parents:
                raw = bytearray(b"TREE" + struct.pack("<BBH", 0, lvl, len(part)))
                raw += b"\x00" * 16                                              # sibling addresses patched below
                raw += struct.pack("<Q", 0)                                      # key 0: the empty string
                for addr, last in part:
                    raw += struct.pack("<QQ", addr, last)
                raw += b"\x00" * (TREE_SIZE - len(raw))
                nodes.append(raw)
            for raw in nodes:
                addrs.append(self.alloc(bytes(raw)))
            for i, addr in enumerate(addrs):                                     # sibling links inside the level
                left = addrs[i - 1] if i > 0 else UNDEF
                right = addrs[i + 1] if i + 1 < len(addrs) else UNDEF
                struct.pack_into("<QQ", self.buf, addr - self.start + 8, left, right)
            # a node's first key is the last name of its left sibling's subtree (libhdf5 walks keys across siblings)
            for i in range(1, len(addrs)):
                struct.pack_into("<Q", self.buf, addrs[i] - self.start + 24, parents[i - 1][0][-1][1])
            level_nodes = [(addr, part[-1][1]) for addr, (part, _) in zip(addrs, parents)]
            if len(level_nodes) == 1:
                break
            level += 1
        btree_addr = level_nodes[0][0]
        header = self.alloc(_object_header([_message(MSG_SYMBOL_TABLE, struct.pack("<QQ", btree_addr, heap_addr))]))
        return header, btree_addr, heap_addr


class File:
    """
    ``File(path, mode)`` with modes ``"r"``, ``"w"`` and ``"a"``; the part of the h5py surface the S^3 I/O layer uses:
    ``keys(group)``, ``has(path)``, ``read(path)``, ``write(group, name, array)``, ``close()``.
    """

    def __init__(self, path: str, mode: str = "r"):
        if mode not in ("r", "w", "a", "r+"):
            raise ValueError(f"unknown mode {mode}")
        self.path, self.mode = path, mode
        self.root: Dict = {}
        self._dirty = False
        exists = os.path.isfile(path)
        if mode == "r":
            self._f = open(path, "rb")
        elif mode == "w" or not exists:
            if mode == "r+":
                raise FileNotFoundError(path)
            self._f = open(path, "w+b")
            self._f.write(b"\x00" * 96)                        # superblock + root entry, filled in by close()
            self._dirty = True
        else:
            self._f = open(path, "r+b")
        if mode != "w" and exists:
            reader = _Reader(self._f)
            self.root = reader.tree()
            self._reader = reader
        self._closed = False

    # ------------------------------------------------------------------ queries
    def _node(self, path: str):
        node = self.root
        for part in [p for p in path.split("/") if p]:
            if not isinstance(node, dict) or part not in node:
                return None
            node = node[part]
        return node

    def keys(self, path: str = "") -> List[str]:
        node = self._node(path)
        return sorted(node.keys(), key=lambda s: s.encode("utf-8")) if isinstance(node, dict) else []

    def has(self, path: str) -> bool:
        return self._node(path) is not None

    def shape(self, path: str) -> tuple:
        node = self._node(path)
        if not isinstance(node, _Dataset):
            raise KeyError(path)
        return node.shape

    def read(self, path: str) -> np.ndarray:
        node = self._node(path)
        if not isinstance(node, _Dataset):
            raise KeyError(path)
        if node.compact is not None:
            raw = node.compact
        elif node.nbytes == 0 or node.address == UNDEF:
            raw = b""
        else:
            self._f.seek(node.address)
            raw = self._f.read(node.nbytes)
        arr = np.frombuffer(raw, dtype=node.dtype, count=node.nbytes // node.dtype.itemsize).reshape(node.shape)
        if node.dtype.byteorder == ">":
            arr = arr.astype(node.dtype.newbyteorder("="))
        return arr.copy()

    # ------------------------------------------------------------------ writing
    def write(self, group: str, name: str, data) -> bool:
        """Create dataset ``group/name``; False (nothing written) if it exists already."""
        if self.mode == "r":
            raise H5Error("file is open read-only")
        node = self.root
        for part in [p for p in group.split("/") if p]:
            nxt = node.setdefault(part, {})
            if not isinstance(nxt, dict):
                raise H5Error(f"{part} is a dataset, not a group")
            node = nxt
        if name in node:
            return False
        arr = np.asarray(data)
        if arr.dtype == np.bool_:
            arr = arr.astype(np.uint8)
        if arr.dtype.byteorder == ">":
            arr = arr.astype(arr.dtype.newbyteorder("<"))
        _encode_dtype(arr.dtype)                                # fail early on unsupported dtypes
        if arr.ndim:                                            # (ascontiguousarray would turn a scalar into [1])
            arr = np.ascontiguousarray(arr)
        self._f.seek(0, os.SEEK_END)
        pad = -self._f.tell() % 8
        if pad:
            self._f.write(b"\x00" * pad)
        address = self._f.tell()
        if arr.nbytes:
            self._f.write(arr.tobytes() if arr.ndim == 0 else memoryview(arr).cast("B"))
        node[name] = _Dataset(tuple(arr.shape), arr.dtype, address, arr.nbytes)
        self._dirty = True
        return True

    def close(self) -> None:
        if self._closed:
            return
        if self.mode != "r" and self._dirty:
            self._f.seek(0, os.SEEK_END)
            pad = -self._f.tell() % 8
            if pad:
                self._f.write(b"\x00" * pad)
            meta = _MetaWriter(self._f.tell())
            header, btree, heap = meta.group(self.root)
            self._f.write(bytes(meta.buf))
            eof = self._f.tell()
            sb = bytearray(SIGNATURE + bytes([0, 0, 0, 0, 0, 8, 8, 0]) + struct.pack("<HHI", LEAF_K, INTERNAL_K, 0))
            sb += struct.pack("<QQQQ", 0, UNDEF, eof, UNDEF)
            sb += struct.pack("<QQII", 0, header, 1, 0) + struct.pack("<QQ", btree, heap)
            assert len(sb) == 96
            self._f.seek(0)
            self._f.write(bytes(sb))
        self._f.close()
        self._closed = True

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()
