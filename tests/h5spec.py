"""Independent structural verifier for HDF5 files, written from the HDF5 File Format Specification (version 1.x objects:
superblock v0, "old style" groups = symbol-table message -> version-1 B-tree of type 0 -> symbol-table nodes + local heap,
version-1 object headers, dataspace v1/v2, datatype v1 classes 0 / 1 / 3, fill value v2, contiguous layout v3, attribute v1).

Test infrastructure only.  It shares no code with `glimslib_b200.backend.minih5` (neither its writer nor its reader): it walks
the raw bytes from the superblock to every dataset and raises `SpecError` on anything the specification forbids -- wrong
signatures or versions, non-zero reserved bytes, misaligned or overlapping structures, unsorted symbol tables, B-tree keys that do
not bracket their children, header sizes that do not add up, addresses beyond the end-of-file address.  What it returns
(`{path: (ndarray, attrs)}`) is compared with what was written by the tests.  Section numbers refer to the specification
("III.A Disk Format: Level 1A1 - Version 1 B-trees" etc.)."""
import struct

import numpy as np

UNDEF = 0xFFFFFFFFFFFFFFFF
SIGNATURE = b"\x89HDF\r\n\x1a\n"


class SpecError(AssertionError):
    pass


def _req(cond, msg):
    if not cond:
        raise SpecError(msg)


class Walker:
    def __init__(self, buf):
        self.b = bytes(buf)
        self.extents = []          # (start, end, what): every structure visited, for the overlap check
        self.superblock()

    # ---- II.A superblock, version 0 --------------------------------------------------------------------------------------
    def superblock(self):
        b = self.b
        _req(b[:8] == SIGNATURE, "format signature")
        _req(b[8] == 0, "superblock version 0")
        _req(b[9] == 0 and b[10] == 0 and b[12] == 0, "free-space / root-entry / shared-header versions are 0")
        _req(b[11] == 0 and b[15] == 0, "reserved bytes of the superblock are 0")
        self.so, self.sl = b[13], b[14]
        _req(self.so == 8 and self.sl == 8, "8-byte offsets and lengths (what libhdf5 writes on 64-bit hosts)")
        self.leaf_k, self.internal_k = struct.unpack_from("<HH", b, 16)
        _req(self.leaf_k > 0 and self.internal_k > 0, "group node K values are positive")
        flags = struct.unpack_from("<I", b, 20)[0]
        _req(flags == 0, "file consistency flags are 0 in a closed file")
        base, free, eof, drv = struct.unpack_from("<QQQQ", b, 24)
        _req(base == 0, "base address 0")
        _req(free == UNDEF, "no free-space info: undefined address")
        _req(drv == UNDEF, "no driver info block: undefined address")
        _req(eof == len(b), "end-of-file address equals the file size")
        self.eof = eof
        # root group symbol table entry (III.C)
        name_off, oh, cache, rsv = struct.unpack_from("<QQII", b, 56)
        _req(name_off == 0 and rsv == 0, "root entry: link name offset 0, reserved 0")
        _req(cache in (0, 1), "root entry cache type 0 or 1")
        self.root_oh = oh
        self.root_cache = struct.unpack_from("<QQ", b, 80) if cache == 1 else None
        self.mark(0, 96, "superblock")

    def mark(self, start, end, what):
        _req(0 <= start < end <= self.eof, "%s [%d, %d) lies inside the file" % (what, start, end))
        self.extents.append((start, end, what))

    def check_no_overlap(self):
        ext = sorted(self.extents)
        for (s0, e0, w0), (s1, e1, w1) in zip(ext, ext[1:]):
            _req(e0 <= s1, "%s [%d,%d) overlaps %s [%d,%d)" % (w0, s0, e0, w1, s1, e1))

    # ---- IV.A.1.a version-1 object header ---------------------------------------------------------------------------------
    def object_header(self, addr):
        b = self.b
        _req(addr % 8 == 0, "object header is 8-byte aligned")
        version, rsv, nmsg, refcnt, size = struct.unpack_from("<BBHII", b, addr)
        _req(version == 1 and rsv == 0, "object header version 1, reserved 0")
        _req(refcnt >= 1, "object reference count >= 1")
        _req(b[addr + 12:addr + 16] == b"\0\0\0\0", "object header prefix is padded to 16 bytes with zeros")
        self.mark(addr, addr + 16 + size, "object header @%d" % addr)
        msgs, blocks, seen = [], [(addr + 16, size)], 0
        while blocks:
            p, left = blocks.pop(0)
            while left >= 8:
                mtype, msize, mflags = struct.unpack_from("<HHB", b, p)
                _req(b[p + 5:p + 8] == b"\0\0\0", "message reserved bytes are 0")
                _req(msize % 8 == 0, "message data size is a multiple of 8 (version-1 headers)")
                _req(8 + msize <= left, "message fits its header block")
                data = b[p + 8:p + 8 + msize]
                if mtype == 0x0010:                       # continuation
                    caddr, clen = struct.unpack_from("<QQ", data, 0)
                    self.mark(caddr, caddr + clen, "continuation block @%d" % caddr)
                    blocks.append((caddr, clen))
                msgs.append((mtype, mflags, data))
                seen += 1
                p += 8 + msize
                left -= 8 + msize
            _req(left == 0, "header block is filled exactly by its messages")
        _req(seen == nmsg, "number of messages matches the header (%d vs %d)" % (seen, nmsg))
        return msgs

    # ---- III.D local heap, III.A B-tree v1 (type 0), III.B symbol-table node ----------------------------------------------
    def heap(self, addr):
        b = self.b
        _req(b[addr:addr + 4] == b"HEAP", "local heap signature")
        _req(b[addr + 4] == 0 and b[addr + 5:addr + 8] == b"\0\0\0", "local heap version 0, reserved 0")
        seg_size, free_head, seg_addr = struct.unpack_from("<QQQ", b, addr + 8)
        self.mark(addr, addr + 32, "local heap header @%d" % addr)
        self.mark(seg_addr, seg_addr + seg_size, "local heap data @%d" % seg_addr)
        _req(seg_size % 8 == 0, "heap data segment size is a multiple of 8")
        # "no free block": the specification's text says the undefined address; libhdf5 itself writes and expects
        # H5HL_FREE_NULL = 1 (H5HLprivate.h), which can never be a real offset (free blocks are 8-byte aligned)
        _req(free_head in (UNDEF, 1) or (free_head % 8 == 0 and free_head + 16 <= seg_size),
             "heap free-list head is 'none' or an aligned offset inside the segment")
        _req(b[seg_addr] == 0, "heap offset 0 holds the empty string (key of the leftmost B-tree entry)")
        seg = b[seg_addr:seg_addr + seg_size]

        def name(off):
            _req(off < seg_size, "name offset inside the heap segment")
            end = seg.index(b"\0", off)
            return seg[off:end].decode("ascii")
        return name

    def btree(self, addr, name, level_expected=None):
        """Returns [(name, object header address)] of the group in B-tree order; checks key bracketing and sibling links."""
        b = self.b
        _req(b[addr:addr + 4] == b"TREE", "B-tree node signature")
        ntype, level, used = struct.unpack_from("<BBH", b, addr + 4)
        _req(ntype == 0, "B-tree node type 0 (group)")
        if level_expected is not None:
            _req(level == level_expected, "child node is one level below its parent")
        k = self.internal_k        # every node of a group B-tree is sized by the internal K; the leaf K sizes symbol-table nodes
        _req(used <= 2 * k, "entries used <= 2K")
        left, right = struct.unpack_from("<QQ", b, addr + 8)
        size = 24 + (2 * k + 1) * 8 + 2 * k * 8
        self.mark(addr, addr + size, "B-tree node @%d" % addr)
        keys = [struct.unpack_from("<Q", b, addr + 24 + 16 * i)[0] for i in range(used + 1)]
        kids = [struct.unpack_from("<Q", b, addr + 24 + 16 * i + 8)[0] for i in range(used)]
        key_names = [name(o) for o in keys]
        _req(key_names == sorted(key_names), "B-tree keys are in ascending name order")
        out = []
        for i, child in enumerate(kids):
            sub = self.btree(child, name, level - 1) if level > 0 else self.snod(child, name)
            _req(len(sub) > 0, "B-tree child is not empty")
            _req(all(key_names[i] < n <= key_names[i + 1] for n, _ in sub),
                 "names of child %d lie in (key[i], key[i+1]]" % i)
            out += sub
        return out, (left, right)

    def snod(self, addr, name):
        b = self.b
        _req(b[addr:addr + 4] == b"SNOD", "symbol table node signature")
        _req(b[addr + 4] == 1 and b[addr + 5] == 0, "symbol table node version 1, reserved 0")
        n = struct.unpack_from("<H", b, addr + 6)[0]
        _req(n <= 2 * self.leaf_k, "symbols in a node <= 2 * leaf K")
        self.mark(addr, addr + 8 + 2 * self.leaf_k * 40, "symbol table node @%d" % addr)
        out = []
        for i in range(n):
            off, oh, cache, rsv = struct.unpack_from("<QQII", b, addr + 8 + 40 * i)
            _req(rsv == 0, "symbol table entry reserved field is 0")
            _req(cache in (0, 1, 2), "symbol table entry cache type")
            out.append((name(off), oh))
        names = [x for x, _ in out]
        _req(names == sorted(names) and len(set(names)) == len(names), "symbols of a node are sorted and unique")
        return out

    # ---- IV.A.2 messages ----------------------------------------------------------------------------------------------------
    @staticmethod
    def dataspace(d):
        version, rank, flags = d[0], d[1], d[2]
        _req(version in (1, 2), "dataspace message version 1 or 2")
        if version == 1:
            _req(d[3] == 0 and d[4:8] == b"\0\0\0\0", "dataspace v1 reserved bytes are 0")
            off = 8
        else:
            _req(d[3] in (0, 1, 2), "dataspace v2 type scalar / simple / null")
            off = 4
        _req(flags & ~1 == 0, "dataspace flags: only 'maximum dimensions present' may be set")
        dims = struct.unpack_from("<%dQ" % rank, d, off)
        need = off + 8 * rank * (2 if flags & 1 else 1)
        _req(need <= len(d) and not any(d[need:]), "dataspace message is padded with zeros")
        return tuple(dims)

    @staticmethod
    def datatype(d):
        cls, version = d[0] & 0x0F, d[0] >> 4
        _req(version == 1, "datatype message version 1")
        bits = d[1] | (d[2] << 8) | (d[3] << 16)
        size = struct.unpack_from("<I", d, 4)[0]
        if cls == 0:                                           # fixed point
            _req(bits & ~0x0F == 0 and bits & 1 == 0, "fixed-point: little-endian, no undefined bits")
            boff, prec = struct.unpack_from("<HH", d, 8)
            _req(boff == 0 and prec == 8 * size, "fixed-point: full precision, bit offset 0")
            return np.dtype("<%s%d" % ("i" if bits & 8 else "u", size))
        if cls == 1:                                           # floating point
            _req(bits & 1 == 0 and (bits >> 4) & 3 == 2, "float: little-endian, mantissa normalisation 'msb implied'")
            boff, prec, eloc, esz, mloc, msz, bias = struct.unpack_from("<HHBBBBI", d, 8)
            ieee = {8: (63, 52, 11, 0, 52, 1023), 4: (31, 23, 8, 0, 23, 127)}[size]
            _req(((bits >> 8) & 0xFF, eloc, esz, mloc, msz, bias) == ieee, "IEEE 754 field layout")
            _req(boff == 0 and prec == 8 * size, "float: full precision")
            return np.dtype("<f%d" % size)
        if cls == 3:                                           # fixed-length string
            _req(bits & 0x0F in (0, 1, 2) and (bits >> 4) & 0x0F in (0, 1), "string padding type and character set")
            return np.dtype("S%d" % size)
        raise SpecError("datatype class %d is outside the subset the reference's files use" % cls)

    def attribute(self, d):
        version, rsv, nsz, tsz, ssz = struct.unpack_from("<BBHHH", d, 0)
        _req(version == 1 and rsv == 0, "attribute message version 1")
        p8 = lambda n: (n + 7) & ~7
        name = d[8:8 + nsz]
        _req(name.endswith(b"\0") and b"\0" not in name[:-1], "attribute name is null-terminated, size includes the terminator")
        o = 8 + p8(nsz)
        dt = self.datatype(d[o:o + tsz])
        o += p8(tsz)
        dims = self.dataspace(d[o:o + ssz])
        o += p8(ssz)
        n = int(np.prod(dims)) if dims else 1
        _req(o + n * dt.itemsize <= len(d), "attribute data fits the message")
        v = np.frombuffer(d[o:o + n * dt.itemsize], dtype=dt).reshape(dims)
        if dt.kind == "S":
            return name[:-1].decode(), v.reshape(-1)[0].decode() if n == 1 else [x.decode() for x in v.reshape(-1)]
        return name[:-1].decode(), (v.reshape(-1)[0] if not dims else v.copy())

    # ---- objects --------------------------------------------------------------------------------------------------------------
    def walk(self):
        out, groups = {}, {}
        self._object(self.root_oh, "", out, groups)
        self.check_no_overlap()
        return out, groups

    def _object(self, oh, path, out, groups):
        msgs = self.object_header(oh)
        attrs = dict(self.attribute(d) for t, _, d in msgs if t == 0x000C)
        sym = [d for t, _, d in msgs if t == 0x0011]
        if sym:
            _req(len(sym) == 1, "one symbol-table message per group")
            bt, hp = struct.unpack_from("<QQ", sym[0], 0)
            if path == "" and self.root_cache:
                _req(self.root_cache == (bt, hp), "root entry scratch-pad caches the B-tree and heap addresses")
            name = self.heap(hp)
            entries, sib = self.btree(bt, name)
            _req(sib == (UNDEF, UNDEF), "root node of a group B-tree has no siblings")
            names = [n for n, _ in entries]
            _req(names == sorted(names) and len(set(names)) == len(names), "group members are sorted and unique")
            groups[path or "/"] = attrs
            for n, child in entries:
                self._object(child, path + "/" + n, out, groups)
            return
        space = [d for t, _, d in msgs if t == 0x0001]
        dtyp = [(f, d) for t, f, d in msgs if t == 0x0003]
        lay = [d for t, _, d in msgs if t == 0x0008]
        _req(len(space) == 1 and len(dtyp) == 1 and len(lay) == 1, "a dataset has one dataspace, datatype and layout message")
        _req(dtyp[0][0] & 1, "datatype message of a dataset is flagged constant")
        for t, _, d in msgs:
            if t == 0x0005:
                _req(d[0] in (1, 2, 3), "fill value message version")
        dims, dt = self.dataspace(space[0]), self.datatype(dtyp[0][1])
        _req(lay[0][0] == 3 and lay[0][1] == 1, "data layout message version 3, contiguous")
        addr, nbytes = struct.unpack_from("<QQ", lay[0], 2)
        n = int(np.prod(dims)) if dims else 1
        _req(nbytes == n * dt.itemsize, "layout size equals element count times element size")
        if nbytes:
            self.mark(addr, addr + nbytes, "raw data of %s" % path)
        out[path] = (np.frombuffer(self.b[addr:addr + nbytes], dtype=dt).reshape(dims).copy(), attrs)


def verify(buf):
    """Walk `buf` (bytes of an HDF5 file); returns ({dataset path: (array, attrs)}, {group path: attrs})."""
    return Walker(buf).walk()
