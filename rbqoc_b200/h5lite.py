"""h5lite - the subset of the HDF5 container the reference's pulse files use, in plain Python + numpy.

The reference stores every optimised pulse and every evaluation result as a flat HDF5 file of named datasets
(`rbqoc.jl:98-139` read side; `spin13.jl:196-233`, `spin.jl:1532-1550`, `spin14.py:222-241` write side), written by
HDF5.jl or h5py with library defaults.  This image has neither libhdf5 nor h5py, so the container is read and written here
from the published file-format specification (HDF5 File Format Specification version 3.0):

reader  superblock 0/1 (symbol-table root group) and 2/3 (object-header root), a user block before the superblock (base
        address), version-1 and version-2 object headers with continuation blocks, groups stored as symbol tables (B-tree
        v1 + local heap + SNOD) or as compact link messages, nested groups, dataspaces (scalar / simple / null),
        fixed-point, IEEE floating-point, fixed-length and variable-length strings (global heap), enumerations over
        integers, compound types, array members; contiguous, compact and chunked (B-tree v1 index; deflate, shuffle,
        fletcher32) layouts.  Dense groups (fractal heap) and the version-4 chunk indices of `libver='latest'` files are
        refused loudly.
writer  the classic layout libhdf5 produces with default properties: superblock 0, version-1 object headers, symbol-table
        groups (leaf K = 4, internal K = 16), contiguous datasets, scalar dataspaces for numbers, fixed-length UTF-8
        strings (what HDF5.jl writes for a `String`).

Pinned against a file libhdf5 itself wrote (scipy's MATLAB 7.3 test fixture, read in `tests/test_h5lite_cpu.py`) and by
write -> read round trips.  Arrays come back with the dimensions stored in the file (row-major, what h5py shows);
`spin.read_save` reverses them into the shapes Julia sees.
"""
from __future__ import annotations

import struct
import zlib

import numpy as np

SIGNATURE = b"\x89HDF\r\n\x1a\n"
UNDEF = 0xFFFFFFFFFFFFFFFF


class H5FormatError(RuntimeError):
    pass


# ------------------------------------------------------------------------------------------------------------- reader

class _Type:
    """A decoded datatype message: numpy dtype for the plain classes, markers for strings."""

    def __init__(self, dtype=None, size=0, kind="plain", base=None, charset="ascii", pad=0, enum=None):
        self.dtype, self.size, self.kind, self.base, self.charset, self.pad, self.enum = dtype, size, kind, base, charset, pad, enum


class File:
    """Read-only view of an HDF5 file: `File(path)[name]` -> numpy array / scalar / str / dict (sub-group)."""

    def __init__(self, source):
        if isinstance(source, (bytes, bytearray, memoryview)):
            self.buf = bytes(source)
        else:
            with open(source, "rb") as fh:
                self.buf = fh.read()
        try:
            self._find_superblock()
            self._links = self._group_links(self.root_header, self.root_btree, self.root_heap)
        except (IndexError, ValueError, struct.error, TypeError) as exc:      # truncated or foreign bytes behind a valid signature
            raise H5FormatError(f"damaged HDF5 file: {exc}") from exc

    # -- low level
    def _u(self, off, n):
        return int.from_bytes(self.buf[off:off + n], "little")

    def _addr(self, off):
        v = self._u(off, self.so)
        return None if v == (1 << (8 * self.so)) - 1 else v + self.base

    def _find_superblock(self):
        off = 0
        while off + 8 <= len(self.buf):
            if self.buf[off:off + 8] == SIGNATURE:
                break
            off = 512 if off == 0 else off * 2            # the superblock sits at 0, 512, 1024, ... (spec II.A)
        else:
            raise H5FormatError("not an HDF5 file (no superblock signature)")
        self.sb_off = off
        ver = self.buf[off + 8]
        self.root_btree = self.root_heap = None
        if ver in (0, 1):
            self.so, self.sl = self.buf[off + 13], self.buf[off + 14]
            self.leaf_k, self.internal_k = self._u(off + 16, 2), self._u(off + 18, 2)
            p = off + 24 + (4 if ver == 1 else 0)
            self.base = 0
            self.base = self._u(p, self.so)
            if self.base == 0 and off != 0:
                self.base = off                            # files with a user block whose base address was left at 0
            p += 4 * self.so                               # base, free-space, end-of-file, driver-info
            # root group symbol-table entry
            self.root_header = self._addr(p + self.so)
            cache = self._u(p + 2 * self.so, 4)
            if cache == 1:
                self.root_btree = self._addr(p + 2 * self.so + 8)
                self.root_heap = self._addr(p + 2 * self.so + 8 + self.so)
        elif ver in (2, 3):
            self.so, self.sl = self.buf[off + 9], self.buf[off + 10]
            self.base = 0
            self.base = self._u(off + 12, self.so)
            if self.base == 0 and off != 0:
                self.base = off
            self.root_header = self._addr(off + 12 + 3 * self.so)
        else:
            raise H5FormatError(f"superblock version {ver} not supported")

    # -- object headers
    def _messages(self, addr):
        """All (type, flags, payload offset, payload size) messages of the object header at `addr`."""
        b = self.buf
        out = []
        if b[addr:addr + 4] == b"OHDR":
            if b[addr + 4] != 2:
                raise H5FormatError("object header version")
            flags = b[addr + 5]
            p = addr + 6
            if flags & 0x20:
                p += 16
            if flags & 0x10:
                p += 4
            n = 1 << (flags & 3)
            size0 = self._u(p, n)
            p += n
            blocks = [(p, size0)]
            track = bool(flags & 0x04)
            while blocks:
                q, size = blocks.pop(0)
                end = q + size
                while q + 4 <= end:
                    mtype, msize, mflags = b[q], self._u(q + 1, 2), b[q + 3]
                    q += 4 + (2 if track else 0)
                    if q + msize > end:
                        break
                    if mtype == 0x10:
                        caddr, clen = self._addr(q), self._u(q + self.so, self.sl)
                        if b[caddr:caddr + 4] != b"OCHK":
                            raise H5FormatError("object header continuation signature")
                        blocks.append((caddr + 4, clen - 8))
                    elif mtype != 0:
                        out.append((mtype, mflags, q, msize))
                    q += msize
            return out
        if b[addr] != 1:
            raise H5FormatError(f"object header version {b[addr]} at {addr}")
        nmsg = self._u(addr + 2, 2)
        size0 = self._u(addr + 8, 4)
        blocks = [(addr + 16, size0)]
        while blocks and len(out) < nmsg + 64:
            q, size = blocks.pop(0)
            end = q + size
            while q + 8 <= end:
                mtype, msize, mflags = self._u(q, 2), self._u(q + 2, 2), b[q + 4]
                q += 8
                if mtype == 0x10:
                    blocks.append((self._addr(q), self._u(q + self.so, self.sl)))
                elif mtype != 0:
                    out.append((mtype, mflags, q, msize))
                q += msize
        return out

    # -- groups
    def _heap_string(self, heap_addr, off):
        b = self.buf
        if b[heap_addr:heap_addr + 4] != b"HEAP":
            raise H5FormatError("local heap signature")
        data = self._addr(heap_addr + 8 + 2 * self.sl)
        s = data + off
        e = b.index(b"\x00", s)
        return b[s:e].decode("utf-8")

    def _btree_group(self, node, heap, links):
        b = self.buf
        if b[node:node + 4] != b"TREE" or b[node + 4] != 0:
            raise H5FormatError("group B-tree node")
        level, used = b[node + 5], self._u(node + 6, 2)
        p = node + 8 + 2 * self.so + self.sl            # first child, after key 0
        for _ in range(used):
            child = self._addr(p)
            p += self.so + self.sl
            if level > 0:
                self._btree_group(child, heap, links)
                continue
            if b[child:child + 4] != b"SNOD":
                raise H5FormatError("symbol table node signature")
            count = self._u(child + 6, 2)
            e = child + 8
            for _ in range(count):
                name = self._heap_string(heap, self._u(e, self.so))
                links[name] = self._addr(e + self.so)
                e += 2 * self.so + 24
        return links

    def _group_links(self, header, btree=None, heap=None):
        links = {}
        if btree is not None and heap is not None:
            return self._btree_group(btree, heap, links)
        for mtype, _, q, size in self._messages(header):
            if mtype == 0x11:                              # symbol table message
                return self._btree_group(self._addr(q), self._addr(q + self.so), links)
            if mtype == 0x02:                              # link info: dense storage?
                flags = self.buf[q + 1]
                p = q + 2 + (8 if flags & 1 else 0)
                if self._addr(p) is not None:
                    raise H5FormatError("dense (fractal-heap) group storage is not supported; rewrite the file with "
                                        "default library versions")
            if mtype == 0x06:                              # link message
                flags = self.buf[q + 1]
                p = q + 2
                ltype = 0
                if flags & 0x08:
                    ltype = self.buf[p]; p += 1
                if flags & 0x04:
                    p += 8
                if flags & 0x10:
                    p += 1
                n = 1 << (flags & 3)
                ln = self._u(p, n); p += n
                name = self.buf[p:p + ln].decode("utf-8"); p += ln
                if ltype == 0:
                    links[name] = self._addr(p)
        return links

    # -- datatypes
    def _datatype(self, q):
        """Decode the datatype message at q; returns (_Type, bytes consumed)."""
        b = self.buf
        cls, ver = b[q] & 0x0F, b[q] >> 4
        bits = self._u(q + 1, 3)
        size = self._u(q + 4, 4)
        p = q + 8
        order = ">" if bits & 1 else "<"
        if cls == 0:                                       # fixed point
            signed = bool(bits & 0x08)
            return _Type(np.dtype(f"{order}{'i' if signed else 'u'}{size}"), size), 12
        if cls == 1:                                       # floating point
            if size not in (2, 4, 8):
                raise H5FormatError(f"{size}-byte floating-point type")
            return _Type(np.dtype(f"{order}f{size}"), size), 20
        if cls == 3:                                       # fixed-length string
            return _Type(np.dtype(f"S{size}"), size, kind="string", pad=bits & 0x0F,
                         charset="utf-8" if (bits >> 4) & 0x0F == 1 else "ascii"), 8
        if cls == 4:                                       # bit field
            return _Type(np.dtype(f"{order}u{size}"), size), 12
        if cls == 9:                                       # variable length
            base, used = self._datatype(p)
            if bits & 0x0F == 1:
                return _Type(None, size, kind="vlen_string", charset="utf-8" if (bits >> 8) & 0x0F == 1 else "ascii"), 8 + used
            return _Type(None, size, kind="vlen", base=base), 8 + used
        if cls == 8:                                       # enumeration
            base, used = self._datatype(p)
            n = bits & 0xFFFF
            p += used
            names = []
            for _ in range(n):
                e = b.index(b"\x00", p)
                names.append(b[p:e].decode("utf-8"))
                ln = e - p + 1
                p += (ln + 7) // 8 * 8 if ver < 3 else ln
            values = np.frombuffer(b, dtype=base.dtype, count=n, offset=p)
            p += n * base.size
            return _Type(base.dtype, size, enum=dict(zip(names, values.tolist()))), p - q
        if cls == 6:                                       # compound
            n = bits & 0xFFFF
            fields = []
            for _ in range(n):
                e = b.index(b"\x00", p)
                name = b[p:e].decode("utf-8")
                ln = e - p + 1
                p += (ln + 7) // 8 * 8 if ver < 3 else ln
                if ver == 1:
                    off = self._u(p, 4)
                    rank = b[p + 4]
                    dims = [self._u(p + 16 + 4 * i, 4) for i in range(rank)]
                    p += 32
                elif ver == 2:
                    off = self._u(p, 4); p += 4
                    dims = []
                else:
                    nb = max(1, (size.bit_length() + 7) // 8)
                    off = self._u(p, nb); p += nb
                    dims = []
                mt, used = self._datatype(p)
                p += used
                if mt.dtype is None:
                    raise H5FormatError("variable-length member of a compound type")
                fields.append((name, off, np.dtype((mt.dtype, tuple(dims))) if dims else mt.dtype))
            dt = np.dtype({"names": [f[0] for f in fields], "offsets": [f[1] for f in fields],
                           "formats": [f[2] for f in fields], "itemsize": size})
            return _Type(dt, size), p - q
        if cls == 10:                                      # array
            rank = b[p]
            p += 1 + (3 if ver < 3 else 0)
            dims = [self._u(p + 4 * i, 4) for i in range(rank)]
            p += 4 * rank + (4 * rank if ver < 3 else 0)
            base, used = self._datatype(p)
            return _Type(np.dtype((base.dtype, tuple(dims))), size), p + used - q
        if cls == 7:                                       # reference
            return _Type(np.dtype(f"<u{size}"), size, kind="reference"), 8
        raise H5FormatError(f"datatype class {cls} not supported")

    # -- data
    def _global_heap_object(self, addr, index):
        b = self.buf
        if b[addr:addr + 4] != b"GCOL":
            raise H5FormatError("global heap signature")
        size = self._u(addr + 8, self.sl)
        p, end = addr + 8 + self.sl, addr + size
        while p + 8 + self.sl <= end:
            idx = self._u(p, 2)
            osize = self._u(p + 8, self.sl)
            if idx == 0:
                break
            if idx == index:
                return b[p + 8 + self.sl:p + 8 + self.sl + osize]
            p += 8 + self.sl + (osize + 7) // 8 * 8
        raise H5FormatError("global heap object not found")

    def _chunks(self, node, rank, out):
        b = self.buf
        if b[node:node + 4] != b"TREE" or b[node + 4] != 1:
            raise H5FormatError("chunk B-tree node")
        level, used = b[node + 5], self._u(node + 6, 2)
        keysize = 8 + 8 * (rank + 1)
        p = node + 8 + 2 * self.so
        for _ in range(used):
            csize, mask = self._u(p, 4), self._u(p + 4, 4)
            offs = [self._u(p + 8 + 8 * i, 8) for i in range(rank)]
            child = self._addr(p + keysize)
            p += keysize + self.so
            if level > 0:
                self._chunks(child, rank, out)
            else:
                out.append((offs, csize, mask, child))
        return out

    def _read_dataset(self, header):
        b = self.buf
        shape, dtype, layout, filters = None, None, None, []
        null_space = False
        for mtype, _, q, size in self._messages(header):
            if mtype == 0x01:
                ver, rank, flags = b[q], b[q + 1], b[q + 2]
                if ver == 1:
                    p = q + 8
                elif ver == 2:
                    null_space = b[q + 3] == 2
                    p = q + 4
                else:
                    raise H5FormatError("dataspace version")
                shape = tuple(self._u(p + self.sl * i, self.sl) for i in range(rank))
            elif mtype == 0x03:
                dtype, _ = self._datatype(q)
            elif mtype == 0x08:
                layout = q
            elif mtype == 0x0B:
                ver, nf = b[q], b[q + 1]
                p = q + (8 if ver == 1 else 2)
                for _ in range(nf):
                    fid = self._u(p, 2)
                    if ver == 1 or fid >= 256:
                        nlen = self._u(p + 2, 2); p += 4
                    else:
                        nlen = 0; p += 2
                    p += 2
                    ncd = self._u(p, 2); p += 2
                    p += (nlen + 7) // 8 * 8 if ver == 1 else nlen
                    cd = [self._u(p + 4 * i, 4) for i in range(ncd)]
                    p += 4 * ncd + (4 if ver == 1 and ncd % 2 else 0)
                    filters.append((fid, cd))
        if shape is None or dtype is None or layout is None:
            raise H5FormatError("object is not a dataset")
        count = int(np.prod(shape)) if shape else 1
        if null_space:
            count = 0
        esize = dtype.size
        ver = b[layout]
        raw = None
        if ver == 3 or ver == 4:
            cls = b[layout + 1]
            if cls == 0:
                n = self._u(layout + 2, 2)
                raw = b[layout + 4:layout + 4 + n]
            elif cls == 1:
                addr = self._addr(layout + 2)
                raw = b"\x00" * (count * esize) if addr is None else b[addr:addr + count * esize]
            elif cls == 2 and ver == 3:
                nd = b[layout + 2]
                btree = self._addr(layout + 3)
                cdims = [self._u(layout + 3 + self.so + 4 * i, 4) for i in range(nd)]
                raw = self._assemble_chunks(btree, shape, cdims[:-1], esize, filters)
            else:
                raise H5FormatError("version-4 chunk indexing (libver='latest') is not supported")
        elif ver in (1, 2):
            nd, cls = b[layout + 1], b[layout + 2]
            p = layout + 8
            addr = None
            if cls != 0:
                addr = self._addr(p); p += self.so
            dims = [self._u(p + 4 * i, 4) for i in range(nd)]
            p += 4 * nd
            if cls == 1:
                raw = b"\x00" * (count * esize) if addr is None else b[addr:addr + count * esize]
            elif cls == 2:
                raw = self._assemble_chunks(addr, shape, dims[:-1], esize, filters)
            else:
                n = self._u(p, 4)
                raw = b[p + 4:p + 4 + n]
        else:
            raise H5FormatError("data layout version")
        return self._decode(raw, dtype, shape, count)

    def _assemble_chunks(self, btree, shape, cdims, esize, filters):
        rank = len(shape)
        out = np.zeros(shape, dtype=np.dtype((np.void, esize)))
        if btree is None:
            return out.tobytes()
        for offs, csize, mask, addr in self._chunks(btree, rank, []):
            data = self.buf[addr:addr + csize]
            for i, (fid, cd) in reversed(list(enumerate(filters))):
                if mask & (1 << i):
                    continue
                if fid == 1:
                    data = zlib.decompress(data)
                elif fid == 2:
                    es = cd[0] if cd else esize
                    n = len(data) // es
                    data = np.frombuffer(data[:n * es], np.uint8).reshape(es, n).T.tobytes() + data[n * es:]
                elif fid == 3:
                    data = data[:-4]
                else:
                    raise H5FormatError(f"filter {fid} not supported")
            chunk = np.frombuffer(data, dtype=out.dtype, count=int(np.prod(cdims))).reshape(cdims)
            sel = tuple(slice(o, min(o + c, s)) for o, c, s in zip(offs, cdims, shape))
            out[sel] = chunk[tuple(slice(0, s.stop - s.start) for s in sel)]
        return out.tobytes()

    def _decode(self, raw, dtype, shape, count):
        if dtype.kind == "vlen_string":
            items = []
            for i in range(count):
                p = i * dtype.size
                n = int.from_bytes(raw[p:p + 4], "little")
                addr = int.from_bytes(raw[p + 4:p + 4 + self.so], "little")
                idx = int.from_bytes(raw[p + 4 + self.so:p + 8 + self.so], "little")
                items.append("" if n == 0 or addr == 0 else self._global_heap_object(addr + self.base, idx)[:n].decode(dtype.charset))
            return items[0] if shape == () else np.array(items, dtype=object).reshape(shape)
        if dtype.kind == "vlen":
            raise H5FormatError("variable-length sequences are not supported")
        arr = np.frombuffer(raw, dtype=dtype.dtype, count=count).reshape(shape if count else (0,) * max(1, len(shape)))
        if dtype.kind == "string":
            def clean(s):
                s = s.split(b"\x00")[0] if dtype.pad in (0, 1) else s.rstrip(b" ")
                return s.decode(dtype.charset)
            if shape == ():
                return clean(arr[()])
            return np.array([clean(s) for s in arr.ravel()], dtype=object).reshape(shape)
        if arr.dtype.byteorder == ">":
            arr = arr.astype(arr.dtype.newbyteorder("="))
        return arr[()] if shape == () else arr.copy()

    # -- mapping interface
    def _is_group(self, header):
        return any(m[0] in (0x11, 0x02, 0x06) for m in self._messages(header)) and \
            not any(m[0] == 0x08 for m in self._messages(header))

    def _load(self, header):
        try:
            if self._is_group(header):
                return {k: self._load(a) for k, a in sorted(self._group_links(header).items())}
            return self._read_dataset(header)
        except (IndexError, struct.error, zlib.error) as exc:
            raise H5FormatError(f"damaged HDF5 object at {header}: {exc}") from exc

    def keys(self):
        return sorted(self._links)

    def __contains__(self, name):
        return name in self._links

    def __iter__(self):
        return iter(self.keys())

    def __getitem__(self, name):
        node, links = None, self._links
        for part in [p for p in name.split("/") if p]:
            if part not in links:
                raise KeyError(name)
            node = links[part]
            links = self._group_links(node) if self._is_group(node) else {}
        if node is None:
            raise KeyError(name)
        return self._load(node)

    def items(self):
        return [(k, self[k]) for k in self.keys()]


def read(path):
    """All datasets of the file as a dict (sub-groups as nested dicts) — the analogue of `read_save` (`rbqoc.jl:129-139`)."""
    f = File(path)
    return {k: v for k, v in f.items()}


# ------------------------------------------------------------------------------------------------------------- writer

def _pad8(b):
    return b + b"\x00" * (-len(b) % 8)


def _dtype_message(dt, strlen=None):
    if strlen is not None:                                 # fixed-length string, null terminated, UTF-8
        return struct.pack("<BBBBI", 0x13, 0x10, 0, 0, strlen)
    dt = np.dtype(dt)
    if dt.kind in "iu":
        bits0 = 0x08 if dt.kind == "i" else 0x00
        return struct.pack("<BBBBIHH", 0x10, bits0, 0, 0, dt.itemsize, 0, 8 * dt.itemsize)
    if dt.kind == "b":
        return struct.pack("<BBBBIHH", 0x10, 0x00, 0, 0, 1, 0, 8)
    if dt.kind == "f" and dt.itemsize == 8:
        return struct.pack("<BBBBIHHBBBBI", 0x11, 0x20, 0x3F, 0, 8, 0, 64, 52, 11, 0, 52, 1023)
    if dt.kind == "f" and dt.itemsize == 4:
        return struct.pack("<BBBBIHHBBBBI", 0x11, 0x20, 0x1F, 0, 4, 0, 32, 23, 8, 0, 23, 127)
    if dt.kind == "c":                                     # h5py's convention for complex: compound {r, i}
        half = dt.itemsize // 2
        m = _dtype_message(np.dtype(f"<f{half}"))
        body = b""
        for name, off in (("r", 0), ("i", half)):
            body += _pad8(name.encode() + b"\x00") + struct.pack("<IB3xII4I", off, 0, 0, 0, 0, 0, 0, 0) + m
        return struct.pack("<BBBBI", 0x16, 2, 0, 0, dt.itemsize) + body
    raise TypeError(f"h5lite cannot store dtype {dt}")


def _message(mtype, body, flags=0):
    body = _pad8(body)
    return struct.pack("<HHB3x", mtype, len(body), flags) + body


class _Image:
    def __init__(self):
        self.b = bytearray()

    def alloc(self, data, align=8):
        self.b += b"\x00" * (-len(self.b) % align)
        addr = len(self.b)
        self.b += data
        return addr


def _object_header(messages):
    body = b"".join(messages)
    return struct.pack("<BBHII4x", 1, 0, len(messages), 1, len(body)) + body


def _dataset_header(img, value):
    """Writes the raw data, returns the object header bytes of a dataset holding `value`."""
    if isinstance(value, str):
        raw = value.encode("utf-8") + b"\x00"
        space = struct.pack("<BBB5x", 1, 0, 0)
        dtype = _dtype_message(None, strlen=len(raw))
    else:
        arr = np.asarray(value)
        if arr.dtype == object:
            raise TypeError("h5lite cannot store object arrays")
        if arr.dtype.kind == "b":
            arr = arr.astype(np.uint8)
        if arr.dtype.byteorder == ">":
            arr = arr.astype(arr.dtype.newbyteorder("<"))
        raw = arr.tobytes(order="C")                    # (np.ascontiguousarray would turn a scalar into shape (1,))
        space = struct.pack("<BBB5x", 1, arr.ndim, 0) + b"".join(struct.pack("<Q", d) for d in arr.shape)
        dtype = _dtype_message(arr.dtype)
    addr = img.alloc(raw) if raw else UNDEF
    layout = struct.pack("<BBQQ", 3, 1, addr, len(raw))
    fill = struct.pack("<BBBBI", 2, 2, 2, 1, 0)           # allocate late, write if set, default fill value
    return _object_header([_message(0x01, space), _message(0x03, dtype, flags=1), _message(0x05, fill),
                           _message(0x08, layout)])


def _write_group(img, entries, leaf_k=4, internal_k=16):
    """entries: name -> value (dict = sub-group).  Returns (object header address, b-tree address, heap address)."""
    names = sorted(entries, key=lambda s: s.encode("utf-8"))
    headers = {}
    for name in names:
        v = entries[name]
        if isinstance(v, dict):
            headers[name] = _write_group(img, v, leaf_k, internal_k)[0]
        else:
            headers[name] = img.alloc(_dataset_header(img, v))
    # local heap: offset 0 holds the empty string (key 0 of the B-tree), then the names, 8-byte aligned
    heap = bytearray(b"\x00" * 8)
    offsets = {}
    for name in names:
        offsets[name] = len(heap)
        heap += _pad8(name.encode("utf-8") + b"\x00")
    free_off = len(heap)
    heap += struct.pack("<QQ", 1, 16)                      # one free block: next = 1 (none), size 16
    heap_data = img.alloc(bytes(heap))
    heap_addr = img.alloc(b"HEAP" + struct.pack("<B3xQQQ", 0, len(heap), free_off, heap_data))
    # symbol table nodes of at most 2*leaf_k entries, allocated at full size as the library does
    per = 2 * leaf_k
    groups = [names[i:i + per] for i in range(0, len(names), per)] or [[]]
    if len(groups) > 2 * internal_k:
        raise ValueError("too many entries for a single-level group B-tree")
    snods, keys = [], [0]
    for g in groups:
        body = b"SNOD" + struct.pack("<BBH", 1, 0, len(g))
        for name in g:
            body += struct.pack("<QQII16x", offsets[name], headers[name], 0, 0)
        body += b"\x00" * (8 + per * 40 - len(body))
        snods.append(img.alloc(body))
        keys.append(offsets[g[-1]] if g else 0)
    used = len(groups) if names else 0
    node = b"TREE" + struct.pack("<BBHQQ", 0, 0, used, UNDEF, UNDEF)
    for i in range(used):
        node += struct.pack("<QQ", keys[i], snods[i])
    node += struct.pack("<Q", keys[used])
    node += b"\x00" * (24 + (2 * internal_k) * 16 + 8 - len(node))
    btree = img.alloc(node)
    header = img.alloc(_object_header([_message(0x11, struct.pack("<QQ", btree, heap_addr))]))
    return header, btree, heap_addr


def write(path, entries, userblock=0):
    """Write `entries` (name -> scalar / numpy array / str / dict of the same) as a classic-format HDF5 file."""
    img = _Image()
    img.b += b"\x00" * 96                                   # superblock, filled in last
    header, btree, heap = _write_group(img, entries)
    eof = len(img.b)
    sb = SIGNATURE + struct.pack("<BBBBBBBBHHI", 0, 0, 0, 0, 0, 8, 8, 0, 4, 16, 0)
    sb += struct.pack("<QQQQ", 0, UNDEF, eof, UNDEF)
    sb += struct.pack("<QQII", 0, header, 1, 0) + struct.pack("<QQ", btree, heap)
    assert len(sb) == 96
    img.b[:96] = sb
    data = bytes(img.b)
    if userblock:
        if userblock < 512 or userblock & (userblock - 1):
            raise ValueError("user block size must be a power of two >= 512")
        # addresses stay relative to the base address, which moves to the end of the user block
        # (the end-of-file address is stored with the base added: the MATLAB 7.3 fixture shows base 512, eof = file size)
        data = b"\x00" * userblock + data[:24] + struct.pack("<Q", userblock) + data[32:40] + \
            struct.pack("<Q", eof + userblock) + data[48:]
    if path is None:
        return data
    with open(path, "wb") as fh:
        fh.write(data)
    return None
