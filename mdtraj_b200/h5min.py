"""Minimal pure-Python reader for the numeric datasets of an MDTraj HDF5 trajectory.

Just enough of the HDF5 file format to pull ``/coordinates`` (and friends) out of files
written by ``mdtraj.formats.HDF5TrajectoryFile`` (``mdtraj/formats/hdf5.py:625-731`` read
semantics: ``/coordinates`` is (n_frames, n_atoms, 3) float32 in nanometres) without
PyTables/h5py, neither of which exists in this image.  Supported: superblock v0/v1,
v1 object headers (+continuations), symbol-table groups (v1 B-tree + local heap),
contiguous and chunked (v1 B-tree) layouts, little-endian fixed-point/IEEE types,
``deflate`` and ``shuffle`` filters.  Anything else raises ``NotImplementedError``.

This feeds the hot path (SURVEY.md section 8(f) "next" #2); it is host-side I/O, not compute.
"""
from __future__ import annotations

import struct
import zlib

import numpy as np

_SIG = b"\x89HDF\r\n\x1a\n"
_UNDEF = 0xFFFFFFFFFFFFFFFF


class H5Min:
    def __init__(self, path):
        with open(path, "rb") as fh:
            self.buf = fh.read()
        b = self.buf
        if b[:8] != _SIG:
            raise ValueError("not an HDF5 file (signature at offset 0 expected)")
        ver = b[8]
        if ver not in (0, 1):
            raise NotImplementedError(f"superblock version {ver}")
        self.so, self.sl = b[13], b[14]  # size of offsets / lengths
        if (self.so, self.sl) != (8, 8):
            raise NotImplementedError("only 8-byte offsets/lengths")
        pos = 24 + (4 if ver == 1 else 0)
        self.base = self._u64(pos)
        pos += 32  # base, free-space, eof, driver
        # root group symbol table entry
        self.root = self._symbol_entry(pos)

    # ---- primitives -----------------------------------------------------
    def _u16(self, p): return struct.unpack_from("<H", self.buf, p)[0]
    def _u32(self, p): return struct.unpack_from("<I", self.buf, p)[0]
    def _u64(self, p): return struct.unpack_from("<Q", self.buf, p)[0]

    def _symbol_entry(self, p):
        name_off, hdr = self._u64(p), self._u64(p + 8)
        cache = self._u32(p + 16)
        btree = heap = None
        if cache == 1:
            btree, heap = self._u64(p + 24), self._u64(p + 32)
        return {"name_off": name_off, "hdr": hdr, "btree": btree, "heap": heap}

    # ---- object headers -------------------------------------------------
    def _messages(self, addr):
        b = self.buf
        if b[addr] != 1:
            raise NotImplementedError("only version-1 object headers")
        nmsg = self._u16(addr + 2)
        size = self._u32(addr + 8)
        blocks = [(addr + 16, size)]
        out = []
        while blocks and len(out) < nmsg:
            p, left = blocks.pop(0)
            end = p + left
            while p + 8 <= end and len(out) < nmsg:
                mtype, msize = self._u16(p), self._u16(p + 2)
                body = p + 8
                if mtype == 0x10:  # continuation
                    blocks.append((self._u64(body), self._u64(body + 8)))
                out.append((mtype, body, msize))
                p = body + msize
        return out

    # ---- groups -----------------------------------------------------------
    def _group_tables(self, entry):
        if entry["btree"] is not None:
            return entry["btree"], entry["heap"]
        for mtype, body, _ in self._messages(entry["hdr"]):
            if mtype == 0x11:
                return self._u64(body), self._u64(body + 8)
        raise NotImplementedError("group without a symbol table (new-style links)")

    def _heap_data(self, heap_addr):
        assert self.buf[heap_addr:heap_addr + 4] == b"HEAP"
        return self._u64(heap_addr + 24)

    def _walk_group_btree(self, node, heap_data, out):
        b = self.buf
        assert b[node:node + 4] == b"TREE"
        level, used = b[node + 5], self._u16(node + 6)
        p = node + 24
        for i in range(used):
            child = self._u64(p + 8)  # key(8) child(8) key(8) ...
            if level > 0:
                self._walk_group_btree(child, heap_data, out)
            else:
                assert b[child:child + 4] == b"SNOD"
                n = self._u16(child + 6)
                q = child + 8
                for _ in range(n):
                    e = self._symbol_entry(q)
                    s = heap_data + e["name_off"]
                    name = b[s:b.index(b"\0", s)].decode()
                    out[name] = e
                    q += 40
            p += 16

    def children(self, entry=None):
        entry = entry or self.root
        btree, heap = self._group_tables(entry)
        out = {}
        self._walk_group_btree(btree, self._heap_data(heap), out)
        return out

    def _lookup(self, path):
        entry = self.root
        for part in [p for p in path.split("/") if p]:
            kids = self.children(entry)
            if part not in kids:
                raise KeyError(path)
            entry = kids[part]
        return entry

    def keys(self):
        return sorted(self.children().keys())

    # ---- datasets ---------------------------------------------------------
    def _dtype(self, body):
        b = self.buf
        cls = b[body] & 0x0F
        bits0 = b[body + 1]
        size = self._u32(body + 4)
        if bits0 & 1:
            raise NotImplementedError("big-endian data")
        if cls == 0:
            signed = bool(bits0 & 0x08)
            return np.dtype(("<i" if signed else "<u") + str(size))
        if cls == 1:
            return np.dtype("<f" + str(size))
        if cls == 3:
            return np.dtype("S" + str(size))
        raise NotImplementedError(f"datatype class {cls}")

    def read(self, path) -> np.ndarray:
        entry = self._lookup(path)
        shape = dtype = layout = None
        filters = []
        for mtype, body, msize in self._messages(entry["hdr"]):
            b = self.buf
            if mtype == 0x01:
                ver, rank = b[body], b[body + 1]
                off = body + (8 if ver == 1 else 4)
                shape = tuple(self._u64(off + 8 * i) for i in range(rank))
            elif mtype == 0x03:
                dtype = self._dtype(body)
            elif mtype == 0x08:
                ver = b[body]
                if ver != 3:
                    raise NotImplementedError(f"data layout version {ver}")
                cls = b[body + 1]
                if cls == 1:
                    layout = ("contiguous", self._u64(body + 2), self._u64(body + 10))
                elif cls == 2:
                    rank = b[body + 2]
                    addr = self._u64(body + 3)
                    dims = tuple(self._u32(body + 11 + 4 * i) for i in range(rank))
                    layout = ("chunked", addr, dims)
                elif cls == 0:
                    sz = self._u16(body + 2)
                    layout = ("compact", body + 4, sz)
            elif mtype == 0x0B:
                ver, nf = b[body], b[body + 1]
                p = body + (8 if ver == 1 else 2)
                for _ in range(nf):
                    fid = self._u16(p)
                    if ver == 1 or fid >= 256:
                        nlen = self._u16(p + 2); p += 2
                    else:
                        nlen = 0
                    ncd = self._u16(p + 4)
                    p += 6
                    p += (nlen + 7) // 8 * 8 if ver == 1 else nlen
                    cd = [self._u32(p + 4 * i) for i in range(ncd)]
                    p += 4 * ncd
                    if ver == 1 and ncd % 2:
                        p += 4
                    filters.append((fid, cd))
        if shape is None or dtype is None or layout is None:
            raise ValueError(f"{path} is not a simple dataset")
        if layout[0] == "contiguous":
            _, addr, size = layout
            if addr == _UNDEF:
                return np.zeros(shape, dtype)
            return np.frombuffer(self.buf, dtype, int(np.prod(shape)), self.base + addr).reshape(shape).copy()
        if layout[0] == "compact":
            _, addr, size = layout
            return np.frombuffer(self.buf, dtype, int(np.prod(shape)), addr).reshape(shape).copy()
        _, addr, dims = layout
        cdims = dims[:-1]
        out = np.zeros(shape, dtype)
        if addr != _UNDEF:
            self._walk_chunk_btree(self.base + addr, len(cdims), cdims, dtype, filters, out)
        return out

    def _walk_chunk_btree(self, node, rank, cdims, dtype, filters, out):
        b = self.buf
        assert b[node:node + 4] == b"TREE" and b[node + 4] == 1
        level, used = b[node + 5], self._u16(node + 6)
        keysize = 8 + 8 * (rank + 1)
        p = node + 24
        for _ in range(used):
            csize, fmask = self._u32(p), self._u32(p + 4)
            offs = tuple(self._u64(p + 8 + 8 * i) for i in range(rank))
            child = self._u64(p + keysize)
            if level > 0:
                self._walk_chunk_btree(self.base + child, rank, cdims, dtype, filters, out)
            else:
                raw = b[self.base + child: self.base + child + csize]
                for i, (fid, cd) in reversed(list(enumerate(filters))):
                    if fmask & (1 << i):
                        continue
                    if fid == 1:
                        raw = zlib.decompress(raw)
                    elif fid == 2:
                        es = cd[0] if cd else dtype.itemsize
                        a = np.frombuffer(raw, np.uint8)
                        n = a.size // es
                        raw = a[: n * es].reshape(es, n).T.tobytes() + a[n * es:].tobytes()
                    else:
                        raise NotImplementedError(f"HDF5 filter {fid}")
                chunk = np.frombuffer(raw, dtype, int(np.prod(cdims))).reshape(cdims)
                sel = tuple(slice(o, min(o + c, s)) for o, c, s in zip(offs, cdims, out.shape))
                sub = tuple(slice(0, s.stop - s.start) for s in sel)
                out[sel] = chunk[sub]
            p += keysize + 8


def load_coordinates(path) -> np.ndarray:
    """``/coordinates`` of an MDTraj HDF5 file as float32 (n_frames, n_atoms, 3), nanometres."""
    xyz = H5Min(path).read("coordinates")
    return np.ascontiguousarray(xyz, dtype=np.float32)


def load_h5(path):
    """An ``mdtraj_b200.Trajectory`` holding the coordinates of an MDTraj HDF5 file (topology not parsed)."""
    from .trajectory import Trajectory
    return Trajectory(load_coordinates(path), None)
