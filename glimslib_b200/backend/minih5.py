"""Minimal HDF5 writer / reader (no libhdf5 / h5py exists in this environment).

Writes real HDF5 files with the oldest, most widely readable structures of the HDF5 File Format Specification:
version-0 superblock, version-1 object headers, "old style" groups (symbol-table message -> v1 B-tree -> one
symbol-table node -> local heap), contiguous little-endian datasets (f64 / f32 / i64 / u64 / i32 / u32 / u8),
simple / scalar dataspaces and version-1 attribute messages (numeric scalars, small numeric arrays, fixed
strings).  That subset is what DOLFIN's ``HDF5File`` / ``XDMFFile`` layouts need (helper_classes.py:1256-1308,
1360-1375).  The reader parses the same subset (any number of B-tree levels / symbol nodes).

The whole tree is kept in memory and serialised on ``close()``; group nodes are sized from the largest group
(the leaf ``K`` of the superblock), so every group is one B-tree node with one symbol-table node.
"""
import struct

import numpy as np

UNDEF = 0xFFFFFFFFFFFFFFFF
SIGNATURE = b"\x89HDF\r\n\x1a\n"


def _pad8(b):
    return b + b"\x00" * ((-len(b)) % 8)


# ---------------------------------------------------------------------------------------------------- datatypes
def _datatype_message(dtype):
    dt = np.dtype(dtype)
    if dt.kind == "f":
        size = dt.itemsize
        if size == 8:
            props = struct.pack("<HHBBBBI", 0, 64, 52, 11, 0, 52, 1023)
            sign = 63
        elif size == 4:
            props = struct.pack("<HHBBBBI", 0, 32, 23, 8, 0, 23, 127)
            sign = 31
        else:
            raise TypeError("unsupported float size")
        head = bytes([0x11, 0x20, sign, 0x00]) + struct.pack("<I", size)      # class 1 v1; LE, implied mantissa msb
        return head + props
    if dt.kind in "iu":
        bits0 = 0x08 if dt.kind == "i" else 0x00                               # bit 3: signed
        head = bytes([0x10, bits0, 0x00, 0x00]) + struct.pack("<I", dt.itemsize)
        return head + struct.pack("<HH", 0, dt.itemsize * 8)
    if dt.kind == "S":
        head = bytes([0x13, 0x00, 0x00, 0x00]) + struct.pack("<I", dt.itemsize)    # class 3 string, null-terminated, ASCII
        return head
    raise TypeError("unsupported dtype %r" % dt)


def _parse_datatype(buf):
    cls = buf[0] & 0x0F
    bits0 = buf[1]
    size = struct.unpack_from("<I", buf, 4)[0]
    if cls == 1:
        return np.dtype("<f%d" % size)
    if cls == 0:
        return np.dtype("<%s%d" % ("i" if bits0 & 0x08 else "u", size))
    if cls == 3:
        return np.dtype("S%d" % size)
    raise TypeError("unsupported HDF5 datatype class %d" % cls)


def _dataspace_message(shape):
    if shape == ():
        return struct.pack("<BBBB4x", 1, 0, 0, 0)
    return struct.pack("<BBBB4x", 1, len(shape), 0, 0) + b"".join(struct.pack("<Q", int(n)) for n in shape)


def _parse_dataspace(buf):
    version, rank = buf[0], buf[1]
    off = 8 if version == 1 else 4
    return tuple(struct.unpack_from("<Q", buf, off + 8 * i)[0] for i in range(rank))


def _message(mtype, data, flags=0):
    data = _pad8(data)
    return struct.pack("<HHB3x", mtype, len(data), flags) + data


def _attribute_message(name, value):
    if isinstance(value, str):
        arr = np.array(value.encode() + b"\x00", dtype="S%d" % (len(value.encode()) + 1))
    else:
        arr = np.asarray(value)
        if arr.dtype.kind == "f":
            arr = arr.astype("<f8")
        elif arr.dtype.kind == "b":
            arr = arr.astype("<i8")
        elif arr.dtype.kind in "iu":
            arr = arr.astype("<i8" if arr.dtype.kind == "i" else "<u8")
    nm = name.encode() + b"\x00"
    dt = _datatype_message(arr.dtype)
    ds = _dataspace_message(arr.shape)
    body = struct.pack("<BxHHH", 1, len(nm), len(dt), len(ds)) + _pad8(nm) + _pad8(dt) + _pad8(ds) + arr.tobytes()
    return _message(0x000C, body)


# ---------------------------------------------------------------------------------------------------- tree model
class Dataset:
    def __init__(self, data):
        self.data = np.ascontiguousarray(data)
        if self.data.dtype.kind == "f" and self.data.dtype.itemsize == 8:
            self.data = self.data.astype("<f8")
        self.attrs = {}


class Group:
    def __init__(self):
        self.children = {}
        self.attrs = {}

    def require_group(self, path):
        g = self
        for part in [p for p in path.split("/") if p]:
            nxt = g.children.get(part)
            if nxt is None:
                nxt = g.children[part] = Group()
            if not isinstance(nxt, Group):
                raise ValueError("%s is a dataset" % part)
            g = nxt
        return g

    def get(self, path):
        g = self
        for part in [p for p in path.split("/") if p]:
            if not isinstance(g, Group) or part not in g.children:
                return None
            g = g.children[part]
        return g

    def create_dataset(self, path, data):
        parts = [p for p in path.split("/") if p]
        g = self.require_group("/".join(parts[:-1]))
        d = Dataset(data)
        g.children[parts[-1]] = d
        return d


# ---------------------------------------------------------------------------------------------------- writer
class _Writer:
    def __init__(self, root):
        self.root = root
        self.buf = bytearray()
        self.leaf_k = max(4, (self._max_children(root) + 1) // 2)
        self.internal_k = 16

    def _max_children(self, g):
        m = len(g.children)
        for c in g.children.values():
            if isinstance(c, Group):
                m = max(m, self._max_children(c))
        return m

    def alloc(self, nbytes, align=8):
        pad = (-len(self.buf)) % align
        self.buf += b"\x00" * pad
        addr = len(self.buf)
        self.buf += b"\x00" * nbytes
        return addr

    def put(self, addr, data):
        self.buf[addr:addr + len(data)] = data

    def write_object_header(self, messages):
        body = b"".join(messages)
        addr = self.alloc(16 + len(body))
        self.put(addr, struct.pack("<BxHII4x", 1, len(messages), 1, len(body)) + body)
        return addr

    def write_dataset(self, d):
        data = d.data
        raw = data.tobytes()
        daddr = self.alloc(len(raw)) if raw else UNDEF
        if raw:
            self.put(daddr, raw)
        msgs = [_message(0x0001, _dataspace_message(data.shape)),
                _message(0x0003, _datatype_message(data.dtype), flags=1),
                _message(0x0005, struct.pack("<BBBB", 2, 2, 2, 0)),
                _message(0x0008, struct.pack("<BBQQ", 3, 1, daddr, len(raw)))]
        msgs += [_attribute_message(k, v) for k, v in d.attrs.items()]
        return self.write_object_header(msgs)

    def write_group(self, g):
        """Returns (object header address, btree address, heap address)."""
        names = sorted(g.children, key=lambda s: s.encode())
        child_addr = {}
        child_scratch = {}
        for nm in names:
            c = g.children[nm]
            if isinstance(c, Group):
                oh, bt, hp = self.write_group(c)
                child_addr[nm], child_scratch[nm] = oh, (1, struct.pack("<QQ", bt, hp))
            else:
                child_addr[nm], child_scratch[nm] = self.write_dataset(c), (0, b"\x00" * 16)
        # local heap: offset 0 holds the empty string
        heap_data = bytearray(b"\x00" * 8)
        name_off = {}
        for nm in names:
            name_off[nm] = len(heap_data)
            heap_data += _pad8(nm.encode() + b"\x00")
        heap_data_addr = self.alloc(len(heap_data))
        self.put(heap_data_addr, bytes(heap_data))
        heap_addr = self.alloc(32)
        self.put(heap_addr, b"HEAP" + struct.pack("<B3xQQQ", 0, len(heap_data), 1, heap_data_addr))
        # one symbol-table node with all entries, sorted by name
        snod_size = 8 + 2 * self.leaf_k * 40
        snod_addr = self.alloc(snod_size)
        ent = b"".join(struct.pack("<QQI4x", name_off[nm], child_addr[nm], child_scratch[nm][0]) + child_scratch[nm][1]
                       for nm in names)
        self.put(snod_addr, b"SNOD" + struct.pack("<BxH", 1, len(names)) + ent)
        # B-tree node (type 0 = group nodes, level 0) with that single child
        bt_size = 24 + (2 * self.internal_k + 1) * 8 + 2 * self.internal_k * 8
        bt_addr = self.alloc(bt_size)
        last_key = name_off[names[-1]] if names else 0
        n_ent = 1 if names else 0
        body = struct.pack("<Q", 0) + (struct.pack("<QQ", snod_addr, last_key) if names else b"")
        self.put(bt_addr, b"TREE" + struct.pack("<BBHQQ", 0, 0, n_ent, UNDEF, UNDEF) + body)
        msgs = [_message(0x0011, struct.pack("<QQ", bt_addr, heap_addr))]
        msgs += [_attribute_message(k, v) for k, v in g.attrs.items()]
        return self.write_object_header(msgs), bt_addr, heap_addr

    def serialise(self):
        self.alloc(96)                                        # superblock placeholder at offset 0
        oh, bt, hp = self.write_group(self.root)
        eof = len(self.buf)
        sb = (SIGNATURE + struct.pack("<BBBBBBBBHHI", 0, 0, 0, 0, 0, 8, 8, 0, self.leaf_k, self.internal_k, 0)
              + struct.pack("<QQQQ", 0, UNDEF, eof, UNDEF)
              + struct.pack("<QQI4x", 0, oh, 1) + struct.pack("<QQ", bt, hp))
        assert len(sb) == 96
        self.put(0, sb)
        return bytes(self.buf)


def write_file(path, root):
    with open(path, "wb") as f:
        f.write(_Writer(root).serialise())


# ---------------------------------------------------------------------------------------------------- reader
class _Reader:
    def __init__(self, buf):
        self.b = buf
        if buf[:8] != SIGNATURE:
            raise IOError("not an HDF5 file")
        if buf[8] != 0 or buf[13] != 8 or buf[14] != 8:
            raise IOError("only version-0 superblocks with 8-byte offsets are supported")
        self.leaf_k, self.internal_k = struct.unpack_from("<HH", buf, 16)
        self.root_oh = struct.unpack_from("<Q", buf, 56 + 8)[0]

    def messages(self, addr):
        version, nmsg, _, size = struct.unpack_from("<BxHII", self.b, addr)
        if version != 1:
            raise IOError("only version-1 object headers are supported")
        out, off, end = [], addr + 16, addr + 16 + size
        while off < end and len(out) < nmsg:
            mtype, msize, flags = struct.unpack_from("<HHB", self.b, off)
            data = self.b[off + 8: off + 8 + msize]
            if mtype == 0x0010:                                # continuation block
                caddr, clen = struct.unpack_from("<QQ", data, 0)
                out += self._messages_raw(caddr, clen)
            else:
                out.append((mtype, data))
            off += 8 + msize
        return out

    def _messages_raw(self, addr, length):
        out, off = [], addr
        while off + 8 <= addr + length:
            mtype, msize, flags = struct.unpack_from("<HHB", self.b, off)
            out.append((mtype, self.b[off + 8: off + 8 + msize]))
            off += 8 + msize
        return out

    def attribute(self, data):
        version, nsz, dsz, ssz = struct.unpack_from("<BxHHH", data, 0)
        off = 8
        name = bytes(data[off:off + nsz]).split(b"\x00")[0].decode()
        off += (nsz + 7) // 8 * 8
        dt = _parse_datatype(data[off:off + dsz])
        off += (dsz + 7) // 8 * 8
        shape = _parse_dataspace(data[off:off + ssz])
        off += (ssz + 7) // 8 * 8
        n = int(np.prod(shape)) if shape else 1
        arr = np.frombuffer(bytes(data[off:off + n * dt.itemsize]), dtype=dt)
        if dt.kind == "S":
            return name, arr[0].split(b"\x00")[0].decode()
        return name, (arr.reshape(shape) if shape else arr[0].item())

    def heap_name(self, heap_addr, off):
        data_addr = struct.unpack_from("<Q", self.b, heap_addr + 24)[0]
        end = self.b.index(b"\x00", data_addr + off)
        return bytes(self.b[data_addr + off:end]).decode()

    def btree_entries(self, bt_addr, heap_addr):
        if self.b[bt_addr:bt_addr + 4] != b"TREE":
            raise IOError("bad B-tree node")
        level, n = struct.unpack_from("<BH", self.b, bt_addr + 5)
        out = []
        for i in range(n):
            child = struct.unpack_from("<Q", self.b, bt_addr + 24 + 8 + 16 * i)[0]
            if level > 0:
                out += self.btree_entries(child, heap_addr)
            else:
                nsym = struct.unpack_from("<H", self.b, child + 6)[0]
                for k in range(nsym):
                    noff, oh = struct.unpack_from("<QQ", self.b, child + 8 + 40 * k)
                    out.append((self.heap_name(heap_addr, noff), oh))
        return out

    def load(self, oh):
        msgs = self.messages(oh)
        attrs = dict(self.attribute(d) for t, d in msgs if t == 0x000C)
        sym = [d for t, d in msgs if t == 0x0011]
        if sym:
            g = Group()
            g.attrs = attrs
            bt, hp = struct.unpack_from("<QQ", sym[0], 0)
            for name, child in self.btree_entries(bt, hp):
                g.children[name] = self.load(child)
            return g
        shape = dt = layout = None
        for t, d in msgs:
            if t == 0x0001:
                shape = _parse_dataspace(d)
            elif t == 0x0003:
                dt = _parse_datatype(d)
            elif t == 0x0008:
                layout = d
        if layout is None or dt is None:
            raise IOError("unsupported object (neither an old-style group nor a dataset)")
        if layout[0] != 3 or layout[1] != 1:
            raise IOError("only contiguous (layout v3 class 1) datasets are supported")
        addr, size = struct.unpack_from("<QQ", layout, 2)
        n = int(np.prod(shape)) if shape else 1
        arr = np.frombuffer(bytes(self.b[addr:addr + size]), dtype=dt)[:n].reshape(shape) if size else np.zeros(shape, dt)
        ds = Dataset(arr.copy())
        ds.attrs = attrs
        return ds


def read_file(path):
    with open(path, "rb") as f:
        buf = f.read()
    r = _Reader(buf)
    return r.load(r.root_oh)
