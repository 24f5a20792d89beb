"""An independent, byte-level walker of the classic HDF5 layout, written from the HDF5 File Format
Specification (version 1.1 structures: superblock 0, version-1 group B-trees, local heaps, symbol
table nodes, version-1 object headers) -- NOT from phyloligo_b200/io_formats.py, whose writer it
checks.  TEST INFRASTRUCTURE ONLY.

h5py / libhdf5 are installed neither in the build image nor on the GPU boxes of this pool (probed:
profiles/r02_env_probe.log), so the files of `--large h5py` (reference bin/phyloligo.py:471-478,
923-930, 787-792; readers bin/phyloligo_comparemat.py:7-14, bin/phyloselect.py:605-622) cannot meet
the real library here.  Two things stand in for it: (1) this walker asserts every field libhdf5
validates when it opens such a file (signatures, versions, sizes of offsets/lengths, K values against
node capacities, heap free list, message sizes and alignment, addresses inside the file, end-of-file
address); (2) the walker itself is pinned to a file that libhdf5 DID write -- the MATLAB 7.3
(= HDF5 with a 512-byte user block) fixture that ships with scipy's test data -- so its reading of
the specification is the library's.
"""
import struct

import numpy as np

SIG = b"\x89HDF\r\n\x1a\n"
UNDEF = 0xFFFFFFFFFFFFFFFF


class SpecError(AssertionError):
    pass


def need(cond, msg):
    if not cond:
        raise SpecError(msg)


class H5File:
    def __init__(self, path):
        self.buf = np.fromfile(path, dtype=np.uint8).tobytes() if not isinstance(path, bytes) else path
        # the superblock sits at 0, 512, 1024, ... (spec III.A: "Location")
        off = 0
        while off < len(self.buf) and self.buf[off:off + 8] != SIG:
            off = 512 if off == 0 else off * 2
        need(self.buf[off:off + 8] == SIG, "no HDF5 signature")
        self.sb = off
        self._superblock()

    # ---- level 0 ----
    def _superblock(self):
        b, o = self.buf, self.sb
        ver, fsver, rootver, r0, shver, so, sl, r1 = struct.unpack_from("<8B", b, o + 8)
        need(ver == 0, "superblock version %d (only 0 handled)" % ver)
        need(fsver == 0 and rootver == 0 and shver == 0 and r0 == 0 and r1 == 0, "superblock version / reserved bytes")
        need(so == 8 and sl == 8, "sizes of offsets / lengths must be 8")
        self.leaf_k, self.internal_k, flags = struct.unpack_from("<HHI", b, o + 16)
        need(self.leaf_k > 0 and self.internal_k > 0, "K values")
        self.base, freespace, self.eof, driver = struct.unpack_from("<QQQQ", b, o + 24)
        need(freespace == UNDEF and driver == UNDEF, "free-space / driver info blocks are not expected")
        need(self.base == self.sb, "base address must be where the superblock is")
        # libhdf5 refuses a file shorter than the stored end-of-file address ("truncated file")
        need(self.eof <= len(self.buf), "file is shorter (%d) than its end-of-file address (%d)" % (len(self.buf), self.eof))
        name_off, oh, cache, rsv = struct.unpack_from("<QQII", b, o + 56)
        need(name_off == 0 and rsv == 0, "root symbol table entry")
        self.root_oh = oh
        self.root_cache = cache
        self.root_scratch = struct.unpack_from("<QQ", b, o + 80) if cache == 1 else None

    def at(self, addr):
        """absolute file position of a (base-relative) address"""
        need(addr != UNDEF, "undefined address dereferenced")
        pos = self.base + addr
        need(pos < len(self.buf), "address 0x%x outside the file" % addr)
        return pos

    # ---- level 1 / 2 ----
    def object_header(self, addr):
        """[(type, flags, data)] of a version-1 object header (spec IV.A.1.a), continuation blocks followed."""
        b, p = self.buf, self.at(addr)
        need(p % 8 == 0, "object header not 8-byte aligned")
        ver, rsv, nmsg, refcount, hsize = struct.unpack_from("<BBHII", b, p)
        need(ver == 1 and rsv == 0, "object header version %d" % ver)
        need(refcount >= 1, "object reference count")
        msgs, blocks = [], [(p + 16, hsize)]  # 12 bytes of prefix, padded to 8-byte alignment
        while blocks:
            pos, size = blocks.pop(0)
            end = pos + size
            need(end <= len(b), "object header block outside the file")
            while pos + 8 <= end and len(msgs) < nmsg:
                mtype, msize, mflags = struct.unpack_from("<HHB", b, pos)
                need(msize % 8 == 0, "message data size %d not a multiple of 8 (type 0x%x)" % (msize, mtype))
                need(pos + 8 + msize <= end, "message overruns its header block")
                data = b[pos + 8:pos + 8 + msize]
                pos += 8 + msize
                msgs.append((mtype, mflags, data))
                if mtype == 0x0010:
                    caddr, clen = struct.unpack_from("<QQ", data, 0)
                    blocks.append((self.at(caddr), clen))
        need(len(msgs) == nmsg, "object header announces %d messages, %d found" % (nmsg, len(msgs)))
        return msgs

    def heap(self, addr):
        b, p = self.buf, self.at(addr)
        need(b[p:p + 4] == b"HEAP" and b[p + 4] == 0 and b[p + 5:p + 8] == b"\0\0\0", "local heap header")
        size, free, data = struct.unpack_from("<QQQ", b, p + 8)
        dp = self.at(data)
        need(dp + size <= len(b), "heap data segment outside the file")
        seg = b[dp:dp + size]
        # free list: (next offset | 1 = end of list, size of this block), blocks >= 16 bytes, inside the segment
        seen = 0
        while free != 1 and free != UNDEF:
            need(free % 8 == 0 and free + 16 <= size, "heap free block at %d outside the segment" % free)
            nxt, fsz = struct.unpack_from("<QQ", seg, free)
            need(fsz >= 16 and free + fsz <= size, "heap free block size")
            free = nxt
            seen += 1
            need(seen < 1000, "heap free list loops")
        return seg

    @staticmethod
    def heap_string(seg, off):
        need(off < len(seg), "name offset outside the heap")
        end = seg.index(b"\0", off)
        return seg[off:end].decode("ascii")

    def group(self, btree_addr, heap_addr):
        """{name: (object header address, cache type, scratch)} of an old-style group."""
        seg = self.heap(heap_addr)
        need(self.heap_string(seg, 0) == "", "heap offset 0 must hold the empty string")
        out = {}

        def node(addr, depth):
            b, p = self.buf, self.at(addr)
            need(b[p:p + 4] == b"TREE", "B-tree node signature")
            ntype, level, used = struct.unpack_from("<BBH", b, p + 4)
            need(ntype == 0, "group B-tree node type")
            need(used <= 2 * self.internal_k, "B-tree node holds %d entries, capacity %d" % (used, 2 * self.internal_k))
            need(p + 24 + (2 * self.internal_k + 1) * 8 + 2 * self.internal_k * 8 <= len(b), "B-tree node (full capacity) outside the file")
            left, right = struct.unpack_from("<QQ", b, p + 8)
            if depth == 0:
                need(left == UNDEF and right == UNDEF, "root B-tree node has siblings")
            keys = [struct.unpack_from("<Q", b, p + 24 + 16 * i)[0] for i in range(used + 1)]
            kids = [struct.unpack_from("<Q", b, p + 32 + 16 * i)[0] for i in range(used)]
            names = [self.heap_string(seg, k) for k in keys]
            need(names == sorted(names), "B-tree keys are not in order")
            for i, kid in enumerate(kids):
                if level > 0:
                    node(kid, depth + 1)
                    continue
                q = self.at(kid)
                need(b[q:q + 4] == b"SNOD" and b[q + 4] == 1 and b[q + 5] == 0, "symbol table node header")
                nsym = struct.unpack_from("<H", b, q + 6)[0]
                need(0 < nsym <= 2 * self.leaf_k, "symbol node holds %d symbols, capacity %d" % (nsym, 2 * self.leaf_k))
                need(q + 8 + 2 * self.leaf_k * 40 <= len(b), "symbol node (full capacity) outside the file")
                last = None
                for s in range(nsym):
                    noff, oh, cache, rsv = struct.unpack_from("<QQII", b, q + 8 + 40 * s)
                    need(rsv == 0 and cache in (0, 1, 2), "symbol table entry")
                    nm = self.heap_string(seg, noff)
                    need(names[i] < nm <= names[i + 1], "symbol %r outside its key interval (%r, %r]" % (nm, names[i], names[i + 1]))
                    need(last is None or last < nm, "symbols of a node are not sorted")
                    last = nm
                    out[nm] = (oh, cache, struct.unpack_from("<QQ", b, q + 8 + 40 * s + 24))
        node(btree_addr, 0)
        return out

    def root_group(self):
        msgs = self.object_header(self.root_oh)
        st = [d for t, f, d in msgs if t == 0x0011]
        need(len(st) == 1, "root object header has no symbol table message")
        btree, heap = struct.unpack_from("<QQ", st[0], 0)
        if self.root_cache == 1:
            need(self.root_scratch == (btree, heap), "root entry scratch pad disagrees with its symbol table message")
        return self.group(btree, heap)

    def dataset(self, oh_addr):
        """dict(shape, dtype, layout, address, size) of a dataset object header; every field checked."""
        info = {}
        for mtype, mflags, data in self.object_header(oh_addr):
            if mtype == 0x0001:  # dataspace, spec IV.A.2.b
                ver, rank, flags = data[0], data[1], data[2]
                need(ver in (1, 2), "dataspace version")
                off = 8 if ver == 1 else 4
                need(len(data) >= off + 8 * rank * (2 if flags & 1 else 1), "dataspace message too short")
                info["shape"] = tuple(struct.unpack_from("<Q", data, off + 8 * i)[0] for i in range(rank))
                if flags & 1:
                    info["maxshape"] = tuple(struct.unpack_from("<Q", data, off + 8 * (rank + i))[0] for i in range(rank))
            elif mtype == 0x0003:  # datatype, spec IV.A.2.d
                cv, b0, b1, b2, size = struct.unpack_from("<BBBBI", data, 0)
                info["type_class"], info["type_version"], info["type_size"] = cv & 15, cv >> 4, size
                need(cv >> 4 in (1, 2, 3), "datatype version")
                if cv & 15 == 1:  # floating point
                    need(b0 & 1 == 0, "big-endian float")
                    need((b0 >> 4) & 3 == 2, "mantissa normalisation must be 'implied msb'")
                    need(b0 & 0x0E == 0, "float padding bits must be zero")
                    bitoff, prec, eloc, esize, mloc, msize, bias = struct.unpack_from("<HHBBBBI", data, 8)
                    ieee = {4: (31, 0, 32, 23, 8, 0, 23, 127), 8: (63, 0, 64, 52, 11, 0, 52, 1023)}
                    need(size in ieee, "float size %d" % size)
                    need((b1, bitoff, prec, eloc, esize, mloc, msize, bias) == ieee[size], "not an IEEE little-endian float%d" % (8 * size))
                    info["dtype"] = np.dtype("<f%d" % size)
            elif mtype == 0x0005:  # fill value, spec IV.A.2.f
                ver = data[0]
                need(ver in (1, 2, 3), "fill value version")
                if ver == 2:
                    alloc, wtime, defined = data[1], data[2], data[3]
                    need(alloc in (1, 2, 3) and wtime in (0, 1, 2) and defined in (0, 1), "fill value fields")
                    if defined:
                        fsz = struct.unpack_from("<I", data, 4)[0]
                        need(8 + fsz <= len(data), "fill value overruns its message")
            elif mtype == 0x0008:  # data layout, spec IV.A.2.i
                ver = data[0]
                need(ver in (1, 2, 3), "layout version %d" % ver)
                info["layout_version"] = ver
                if ver == 3:
                    cls = data[1]
                    info["layout"] = {0: "compact", 1: "contiguous", 2: "chunked"}.get(cls)
                    need(info["layout"] is not None, "layout class")
                    if cls == 1:
                        info["address"], info["size"] = struct.unpack_from("<QQ", data, 2)
                    elif cls == 0:
                        sz = struct.unpack_from("<H", data, 2)[0]
                        info["compact"] = data[4:4 + sz]
                else:  # versions 1 and 2 (libhdf5 <= 1.6): dimensionality, class, reserved, [address], 32-bit dims
                    ndim, cls = data[1], data[2]
                    info["layout"] = {0: "compact", 1: "contiguous", 2: "chunked"}.get(cls)
                    need(info["layout"] is not None, "layout class")
                    pos = 8
                    if cls != 0:
                        info["address"] = struct.unpack_from("<Q", data, pos)[0]
                        pos += 8
                    dims = struct.unpack_from("<%dI" % ndim, data, pos)
                    pos += 4 * ndim
                    if cls == 1:
                        info["size"] = int(np.prod(dims, dtype=np.int64))  # element size is the last "dimension"
                    elif cls == 0:
                        sz = struct.unpack_from("<I", data, pos)[0]
                        info["compact"] = data[pos + 4:pos + 4 + sz]
        return info

    def read(self, oh_addr):
        d = self.dataset(oh_addr)
        need("shape" in d and "dtype" in d and d.get("layout") == "contiguous", "not a contiguous float dataset: %r" % (d,))
        n = int(np.prod(d["shape"], dtype=np.int64))
        need(d["size"] == n * d["dtype"].itemsize, "layout size %d != %d elements x %d bytes" % (d["size"], n, d["dtype"].itemsize))
        if n == 0:
            return np.zeros(d["shape"], d["dtype"])
        p = self.at(d["address"])
        need(p + d["size"] <= self.eof, "raw data beyond the end-of-file address")
        return np.frombuffer(self.buf, dtype=d["dtype"], count=n, offset=p).reshape(d["shape"])
