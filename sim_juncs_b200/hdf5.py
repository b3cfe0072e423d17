"""A small HDF5 writer and reader with no libhdf5.

The reference writes `field_samples.h5` through H5Cpp (src/disp.cpp:758-923): nested groups, 1-D datasets
of native double / native hsize and of three compound types ({Re,Im}, {x,y,z}, source_info), no attributes,
no chunking, no filters.  Neither libhdf5 nor h5py exists in this image, so this module writes that subset
of the format byte by byte, in the encoding libhdf5 1.8 uses with default property lists:

  superblock v0 | v1 object headers | symbol-table groups (local heap + v1 B-tree + SNOD nodes)
  dataspace v1 | datatype v1 (fixed point, IEEE float, compound) | fill value | contiguous layout v3

`H5Reader` is an independent decoder of the same structures (plus the older layout v1/v2, datatype v2/v3,
header continuation blocks and multi-level B-trees) with an h5py-like access pattern, so the reference's
analysis scripts' reads (scripts/phases.py:684-733, scripts/time_space.py:17-49) can be replayed on our files.
The reader is pinned on a file written by the real library that ships inside scipy's test data
(scipy/io/matlab/tests/data/testhdf5_7.4_GLNX86.mat, an HDF5 file behind a 512-byte MATLAB header); the writer's
float64 datatype message, symbol-table message, heap free list and B-tree/SNOD framing are byte-compared with
that file in tests/test_hdf5.py.
"""
import struct

import numpy as np

UNDEF = 0xFFFFFFFFFFFFFFFF
SIGNATURE = b"\x89HDF\r\n\x1a\n"
INTERNAL_K = 16          # libhdf5 default: a group B-tree node holds up to 2*16 children
LEAF_K_DEFAULT = 4       # libhdf5 default: a symbol node holds up to 2*4 entries
HEAP_FREE_NULL = 1       # H5HL_FREE_NULL: end of the local heap's free list


def _pad8(b):
    return b + b"\0" * (-len(b) % 8)


# ------------------------------------------------------------------------------------------------ datatypes
def encode_dtype(dt):
    """Datatype message body (version 1) of a numpy dtype: little-endian ints, IEEE floats, compounds of those."""
    dt = np.dtype(dt)
    if dt.names:
        out = struct.pack("<BBBBI", 0x16, len(dt.names) & 0xFF, (len(dt.names) >> 8) & 0xFF, 0, dt.itemsize)
        for name in dt.names:
            sub, off = dt.fields[name][0], dt.fields[name][1]
            nm = name.encode() + b"\0"
            out += nm + b"\0" * (-len(nm) % 8)
            out += struct.pack("<IB3xI4x4I", off, 0, 0, 0, 0, 0, 0)   # offset, rank 0, permutation, 4 dim sizes
            out += encode_dtype(sub)
        return out
    if dt.byteorder == ">":
        raise ValueError("big-endian arrays are not written")
    if dt.kind == "f":
        props = {4: (0, 32, 23, 8, 0, 23, 127), 8: (0, 64, 52, 11, 0, 52, 1023)}[dt.itemsize]
        return struct.pack("<BBBBI", 0x11, 0x20, 8 * dt.itemsize - 1, 0, dt.itemsize) + struct.pack("<HHBBBBI", *props)
    if dt.kind in "iu":
        return struct.pack("<BBBBI", 0x10, 0x08 if dt.kind == "i" else 0x00, 0, 0, dt.itemsize) + \
            struct.pack("<HH", 0, 8 * dt.itemsize)
    raise ValueError("unsupported dtype %r" % (dt,))


def decode_dtype(buf, pos):
    """-> (numpy dtype, position after the message)."""
    cv, b0, b1, b2, size = struct.unpack_from("<BBBBI", buf, pos)
    cls, ver = cv & 0x0F, cv >> 4
    pos += 8
    if cls == 0:
        order = ">" if b0 & 1 else "<"
        return np.dtype("%s%s%d" % (order, "i" if b0 & 8 else "u", size)), pos + 4
    if cls == 1:
        order = ">" if b0 & 1 else "<"
        return np.dtype("%sf%d" % (order, size)), pos + 12
    if cls == 3:
        return np.dtype("S%d" % size), pos
    if cls == 6:
        n = b0 | (b1 << 8)
        names, formats, offsets = [], [], []
        for _ in range(n):
            end = buf.index(b"\0", pos)
            name = bytes(buf[pos:end]).decode()
            if ver >= 3:
                pos = end + 1
                nb = max(1, (int(size).bit_length() + 7) // 8)
                off = int.from_bytes(buf[pos:pos + nb], "little")
                pos += nb
            else:
                pos += ((end - pos) + 8) // 8 * 8
                off = struct.unpack_from("<I", buf, pos)[0]
                pos += 4
                if ver == 1:
                    pos += 28
            sub, pos = decode_dtype(buf, pos)
            names.append(name)
            formats.append(sub)
            offsets.append(off)
        return np.dtype({"names": names, "formats": formats, "offsets": offsets, "itemsize": size}), pos
    raise ValueError("datatype class %d is not supported" % cls)


# ------------------------------------------------------------------------------------------------ writer
class _Group:
    def __init__(self):
        self.children = {}


class H5Writer:
    """Collects groups and datasets, then writes the whole file in one pass.

        w = H5Writer(path); w.create_group("info"); w.create_dataset("info/time_bounds", array); w.close()
    """

    def __init__(self, path):
        self.path = path
        self.root = _Group()
        self.closed = False

    def _walk(self, path, create):
        g = self.root
        for part in [p for p in path.split("/") if p]:
            if part not in g.children:
                if not create:
                    raise KeyError(path)
                g.children[part] = _Group()
            g = g.children[part]
            if not isinstance(g, _Group):
                raise ValueError("%s is a dataset" % part)
        return g

    def create_group(self, path):
        self._walk(path, True)

    def create_dataset(self, path, data):
        parts = [p for p in path.split("/") if p]
        g = self._walk("/".join(parts[:-1]), True)
        if parts[-1] in g.children:
            raise ValueError("%s exists" % path)
        a = np.ascontiguousarray(data)
        if a.dtype.byteorder == ">":
            a = a.astype(a.dtype.newbyteorder("<"))
        g.children[parts[-1]] = a

    # ---- serialisation
    def _max_entries(self, g):
        m = len(g.children)
        for c in g.children.values():
            if isinstance(c, _Group):
                m = max(m, self._max_entries(c))
        return m

    def _alloc(self, blob):
        self.buf += b"\0" * (-len(self.buf) % 8)
        addr = len(self.buf)
        self.buf += blob
        return addr

    @staticmethod
    def _message(mtype, body, flags=0):
        body = _pad8(body)
        return struct.pack("<HHB3x", mtype, len(body), flags) + body

    def _object_header(self, messages):
        body = b"".join(messages)
        return self._alloc(struct.pack("<BBHII4x", 1, 0, len(messages), 1, len(body)) + body)

    def _emit_dataset(self, a):
        if a.dtype.names and sum(a.dtype.fields[n][0].itemsize for n in a.dtype.names) != a.dtype.itemsize:
            flat = np.zeros(a.size * a.dtype.itemsize, dtype=np.uint8)    # padding bytes of a struct are written as zeros
            view = flat.view(a.dtype).reshape(a.shape)
            for name in a.dtype.names:
                view[name] = a[name]
            raw = flat.tobytes()
        else:
            raw = a.tobytes()
        addr = self._alloc(raw) if raw else UNDEF
        shape = a.shape if a.ndim else (1,)
        space = struct.pack("<BBB5x", 1, len(shape), 0) + b"".join(struct.pack("<Q", d) for d in shape)
        msgs = [self._message(0x0001, space),
                self._message(0x0003, encode_dtype(a.dtype), flags=1),
                self._message(0x0005, struct.pack("<BBBBI", 1, 2, 2, 1, 0), flags=1),   # late alloc, fill if set, default value
                self._message(0x0008, struct.pack("<BBQQ", 3, 1, addr, len(raw)))]
        return self._object_header(msgs)

    def _emit_group(self, g):
        """-> (object header address, B-tree address, heap address)"""
        names = sorted(g.children, key=lambda s: s.encode())
        child = {}
        for nm in names:
            c = g.children[nm]
            child[nm] = self._emit_group(c) if isinstance(c, _Group) else (self._emit_dataset(c), None, None)
        # local heap: offset 0 holds the empty string (key 0 of the B-tree), names follow, one free block closes it
        data = bytearray(8)
        name_off = {}
        for nm in names:
            name_off[nm] = len(data)
            data += _pad8(nm.encode() + b"\0")
        free_off = len(data)
        data += struct.pack("<QQ", HEAP_FREE_NULL, 16)
        heap_addr = self._alloc(b"")
        data_addr = heap_addr + 32
        self._alloc(b"HEAP" + struct.pack("<B3xQQQ", 0, len(data), free_off, data_addr) + bytes(data))
        # symbol nodes
        per = 2 * self.leaf_k
        node_addrs, last_names = [], []
        for s in range(0, len(names), per):
            part = names[s:s + per]
            blob = b"SNOD" + struct.pack("<BBH", 1, 0, len(part))
            for nm in part:
                hdr, bt, hp = child[nm]
                if bt is not None:
                    blob += struct.pack("<QQI4xQQ", name_off[nm], hdr, 1, bt, hp)
                else:
                    blob += struct.pack("<QQI4x16x", name_off[nm], hdr, 0)
            blob += b"\0" * (8 + per * 40 - len(blob))
            node_addrs.append(self._alloc(blob))
            last_names.append(part[-1])
        if len(node_addrs) > 2 * INTERNAL_K:
            raise AssertionError("leaf_k too small")
        tree = b"TREE" + struct.pack("<BBHQQ", 0, 0, len(node_addrs), UNDEF, UNDEF) + struct.pack("<Q", 0)
        for addr, last in zip(node_addrs, last_names):
            tree += struct.pack("<QQ", addr, name_off[last])
        tree += b"\0" * (24 + (4 * INTERNAL_K + 1) * 8 - len(tree))
        bt_addr = self._alloc(tree)
        hdr_addr = self._object_header([self._message(0x0011, struct.pack("<QQ", bt_addr, heap_addr))])
        return hdr_addr, bt_addr, heap_addr

    def close(self):
        if self.closed:
            return
        self.closed = True
        self.leaf_k = max(LEAF_K_DEFAULT, -(-self._max_entries(self.root) // (4 * INTERNAL_K)))
        if self.leaf_k > 0xFFFF:
            raise ValueError("too many entries in one group")
        self.buf = bytearray(96)                      # superblock v0 with 8-byte offsets and lengths
        hdr, bt, hp = self._emit_group(self.root)
        self.buf += b"\0" * (-len(self.buf) % 8)
        sb = SIGNATURE + struct.pack("<BBBBBBBB", 0, 0, 0, 0, 0, 8, 8, 0) + struct.pack("<HHI", self.leaf_k, INTERNAL_K, 0)
        sb += struct.pack("<QQQQ", 0, UNDEF, len(self.buf), UNDEF)
        sb += struct.pack("<QQI4xQQ", 0, hdr, 1, bt, hp)
        assert len(sb) == 96
        self.buf[0:96] = sb
        with open(self.path, "wb") as fp:
            fp.write(self.buf)

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        if exc[0] is None:
            self.close()


# ------------------------------------------------------------------------------------------------ reader
class H5Dataset:
    def __init__(self, reader, shape, dtype, addr, size, inline=None):
        self._r, self.shape, self.dtype, self._addr, self._size, self._inline = reader, tuple(shape), dtype, addr, size, inline

    def read(self):
        n = int(np.prod(self.shape)) if self.shape else 1
        if n == 0 or (self._addr == UNDEF and self._inline is None):
            return np.zeros(self.shape, dtype=self.dtype)
        raw = self._inline if self._inline is not None else self._r.buf[self._r.base + self._addr:
                                                                         self._r.base + self._addr + n * self.dtype.itemsize]
        return np.frombuffer(raw, dtype=self.dtype, count=n).reshape(self.shape).copy()

    def __getitem__(self, key):
        return self.read()[key]

    def __len__(self):
        return self.shape[0] if self.shape else 1

    def __array__(self, dtype=None, copy=None):
        a = self.read()
        return a if dtype is None else a.astype(dtype)


class H5Group:
    def __init__(self, reader, btree, heap):
        self._r, self._btree, self._heap = reader, btree, heap
        self._entries = None

    def _load(self):
        if self._entries is None:
            self._entries = self._r._group_entries(self._btree, self._heap)
        return self._entries

    def keys(self):
        return list(self._load().keys())

    def __iter__(self):
        return iter(self.keys())

    def __len__(self):
        return len(self._load())

    def __contains__(self, k):
        return k in self._load()

    def __getitem__(self, path):
        obj = self
        for part in [p for p in path.split("/") if p]:
            if not isinstance(obj, H5Group):
                raise KeyError(path)
            obj = obj._r._object(obj._load()[part])
        return obj


class H5Reader(H5Group):
    """h5py-like read access: f["info"]["sources"]["wavelen"][0], f.keys() in name order, len(), np.array(ds)."""

    def __init__(self, path):
        with open(path, "rb") as fp:
            self.buf = fp.read()
        off = 0
        while self.buf[off:off + 8] != SIGNATURE:
            off = 512 if off == 0 else off * 2
            if off >= len(self.buf):
                raise ValueError("not an HDF5 file")
        self.sb_off = off
        ver = self.buf[off + 8]
        if ver > 1:
            raise ValueError("superblock version %d is not supported" % ver)
        so, sl = self.buf[off + 13], self.buf[off + 14]
        if (so, sl) != (8, 8):
            raise ValueError("only 8-byte offsets and lengths are supported")
        self.leaf_k, self.internal_k = struct.unpack_from("<HH", self.buf, off + 16)
        p = off + 24 + (4 if ver == 1 else 0)
        self.base, _, self.eof, _ = struct.unpack_from("<QQQQ", self.buf, p)
        p += 32
        _, hdr, cache, bt, hp = struct.unpack_from("<QQI4xQQ", self.buf, p)
        if self.base == 0 and off != 0 and ver == 0:
            self.base = 0
        self.root_header = hdr
        root = self._object(hdr)
        super().__init__(self, root._btree, root._heap)

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        return False

    def _messages(self, addr):
        """(type, flags, body offset, body size) of every message of a version-1 object header."""
        p = self.base + addr
        ver, _, nmsg, _, size = struct.unpack_from("<BBHII", self.buf, p)
        if ver != 1:
            raise ValueError("object header version %d is not supported" % ver)
        blocks = [(p + 16, size)]
        out = []
        while blocks:
            q, sz = blocks.pop(0)
            end = q + sz
            while q + 8 <= end and len(out) < nmsg:
                mtype, msize, flags = struct.unpack_from("<HHB", self.buf, q)
                out.append((mtype, flags, q + 8, msize))
                if mtype == 0x0010:
                    caddr, clen = struct.unpack_from("<QQ", self.buf, q + 8)
                    blocks.append((self.base + caddr, clen))
                q += 8 + msize
        return out

    def _object(self, addr):
        msgs = self._messages(addr)
        kinds = {m[0]: m for m in msgs}
        if 0x0011 in kinds:
            bt, hp = struct.unpack_from("<QQ", self.buf, kinds[0x0011][2])
            return H5Group(self, bt, hp)
        if 0x0001 not in kinds or 0x0003 not in kinds or 0x0008 not in kinds:
            raise ValueError("object at %d is neither a symbol-table group nor a dataset" % addr)
        q = kinds[0x0001][2]
        sver, rank, sflags = struct.unpack_from("<BBB", self.buf, q)
        q += 8 if sver == 1 else 4
        shape = struct.unpack_from("<%dQ" % rank, self.buf, q)
        dtype, _ = decode_dtype(self.buf, kinds[0x0003][2])
        q = kinds[0x0008][2]
        lver = self.buf[q]
        if lver == 3:
            lclass = self.buf[q + 1]
            if lclass == 1:
                a, s = struct.unpack_from("<QQ", self.buf, q + 2)
                return H5Dataset(self, shape, dtype, a, s)
            if lclass == 0:
                s = struct.unpack_from("<H", self.buf, q + 2)[0]
                return H5Dataset(self, shape, dtype, UNDEF, s, inline=self.buf[q + 4:q + 4 + s])
            raise ValueError("chunked datasets are not supported")
        ndim, lclass = self.buf[q + 1], self.buf[q + 2]
        if lclass != 1:
            raise ValueError("only contiguous storage is supported for layout versions 1 and 2")
        a = struct.unpack_from("<Q", self.buf, q + 8)[0]
        return H5Dataset(self, shape, dtype, a, 0)

    def _heap_string(self, heap_addr, off):
        p = self.base + heap_addr
        if self.buf[p:p + 4] != b"HEAP":
            raise ValueError("bad local heap signature")
        data_addr = struct.unpack_from("<Q", self.buf, p + 24)[0]
        s = self.base + data_addr + off
        return bytes(self.buf[s:self.buf.index(b"\0", s)]).decode()

    def _group_entries(self, btree, heap):
        out = {}

        def node(addr):
            p = self.base + addr
            if self.buf[p:p + 4] == b"TREE":
                ntype, level, used = struct.unpack_from("<BBH", self.buf, p + 4)
                if ntype != 0:
                    raise ValueError("not a group B-tree")
                q = p + 24
                for i in range(used):
                    node(struct.unpack_from("<Q", self.buf, q + 8 + 16 * i)[0])
            elif self.buf[p:p + 4] == b"SNOD":
                n = struct.unpack_from("<H", self.buf, p + 6)[0]
                for i in range(n):
                    noff, hdr = struct.unpack_from("<QQ", self.buf, p + 8 + 40 * i)
                    out[self._heap_string(heap, noff)] = hdr
            else:
                raise ValueError("bad group node signature at %d" % addr)

        node(btree)
        return out


def File(path, mode="r"):
    """h5py.File look-alike for the read patterns of the reference's scripts."""
    if mode != "r":
        raise ValueError("read-only")
    return H5Reader(path)
