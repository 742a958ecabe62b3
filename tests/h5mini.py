"""Minimal reader for classic-layout HDF5 files (superblock v0/v1, v1 object headers, symbol-table groups, contiguous or
compact datasets of f64 / integers / strings / complex numbers ({r, i} compounds), variable-length strings in global heaps, v1-v3 attributes).

Test infrastructure: there is no h5py / libhdf5 in this image, so the output writer of the host program
(mc_old_b200/host/h5lite.cpp) is checked by (i) reading the reference's own committed output.h5 files — written by the
real HDF5 library — with this reader, which validates the reader, and (ii) reading our files with the same reader.
Follows the HDF5 File Format Specification; deliberately strict (raises on anything it does not understand)."""
import struct

import numpy as np

UNDEF = 0xFFFFFFFFFFFFFFFF


class H5Error(Exception):
    pass


class Node:
    def __init__(self, kind):
        self.kind = kind          # "group" | "dataset"
        self.children = {}        # groups
        self.attrs = {}
        self.value = None         # datasets: numpy array / scalar / str
        self.shape = None
        self.dtype = None

    def __getitem__(self, path):
        n = self
        for part in [p for p in path.split("/") if p]:
            n = n.children[part]
        return n

    def walk(self, prefix=""):
        for k in sorted(self.children):
            c = self.children[k]
            yield prefix + "/" + k, c
            if c.kind == "group":
                yield from c.walk(prefix + "/" + k)


class File:
    def __init__(self, path):
        with open(path, "rb") as f:
            self.b = f.read()
        b = self.b
        if b[:8] != b"\x89HDF\r\n\x1a\n":
            raise H5Error("not an HDF5 file")
        ver = b[8]
        if ver not in (0, 1):
            raise H5Error("superblock version %d not supported" % ver)
        self.so, self.sl = b[13], b[14]
        if (self.so, self.sl) != (8, 8):
            raise H5Error("only 8-byte offsets/lengths")
        self.leaf_k, self.internal_k = struct.unpack_from("<HH", b, 16)
        p = 24 if ver == 0 else 28
        self.base, _free, self.eof, _drv = struct.unpack_from("<QQQQ", b, p)
        if self.eof > len(b):
            raise H5Error("end-of-file address %d beyond the file (%d bytes)" % (self.eof, len(b)))
        root = p + 32
        _name_off, header = struct.unpack_from("<QQ", b, root)
        self._gcol = {}
        self.root = self._object(header)

    # ---- low level ----
    def _messages(self, addr):
        b = self.b
        ver, _r, nmsg, _ref, size = struct.unpack_from("<BBHII", b, addr)
        if ver != 1:
            raise H5Error("object header version %d at %d" % (ver, addr))
        out = []
        blocks = [(addr + 16, size)]
        while blocks:
            p, n = blocks.pop(0)
            end = p + n
            while p + 8 <= end and len(out) < nmsg:
                mtype, msize, flags = struct.unpack_from("<HHB", b, p)
                data = b[p + 8:p + 8 + msize]
                p += 8 + msize
                if mtype == 0x0010:
                    off, ln = struct.unpack_from("<QQ", data, 0)
                    blocks.append((off, ln))
                out.append((mtype, data, flags))
        return out

    def _heap_string(self, heap_addr, off):
        b = self.b
        if b[heap_addr:heap_addr + 4] != b"HEAP":
            raise H5Error("bad local heap at %d" % heap_addr)
        dsize, _free, daddr = struct.unpack_from("<QQQ", b, heap_addr + 8)
        if off >= dsize:
            raise H5Error("heap offset outside the data segment")
        end = b.index(b"\0", daddr + off)
        return b[daddr + off:end].decode()

    def _symbols(self, btree, heap):
        b = self.b
        if b[btree:btree + 4] != b"TREE":
            raise H5Error("bad B-tree node at %d" % btree)
        ntype, level, used = struct.unpack_from("<BBH", b, btree + 4)
        if ntype != 0:
            raise H5Error("not a group B-tree")
        out = []
        for i in range(used):
            child = struct.unpack_from("<Q", b, btree + 24 + 8 + 16 * i)[0]
            if level > 0:
                out += self._symbols(child, heap)
                continue
            if b[child:child + 4] != b"SNOD":
                raise H5Error("bad symbol node at %d" % child)
            nsym = struct.unpack_from("<H", b, child + 6)[0]
            if nsym > 2 * self.leaf_k:
                raise H5Error("symbol node holds %d > 2K entries" % nsym)
            for s in range(nsym):
                name_off, header = struct.unpack_from("<QQ", b, child + 8 + 40 * s)
                out.append((self._heap_string(heap, name_off), header))
        names = [n for n, _ in out]
        if names != sorted(names):
            raise H5Error("group entries are not sorted: %r" % names)
        return out

    def _global_heap_object(self, addr, idx):
        b = self.b
        if addr not in self._gcol:
            if b[addr:addr + 4] != b"GCOL":
                raise H5Error("bad global heap collection at %d" % addr)
            size = struct.unpack_from("<Q", b, addr + 8)[0]
            if size < 4096 or addr + size > len(b):
                raise H5Error("global heap collection size %d" % size)
            objs, p = {}, addr + 16
            while p + 16 <= addr + size:
                i, _ref, _r, osize = struct.unpack_from("<HHIQ", b, p)
                if i == 0:
                    break
                objs[i] = b[p + 16:p + 16 + osize]
                p += 16 + ((osize + 7) & ~7)
            self._gcol[addr] = objs
        return self._gcol[addr][idx]

    @staticmethod
    def _dataspace(data):
        ver, rank, flags = data[0], data[1], data[2]
        if ver == 1:
            p = 8
        elif ver == 2:
            p = 4
            if data[3] == 2:
                return None  # null dataspace
        else:
            raise H5Error("dataspace version %d" % ver)
        return tuple(struct.unpack_from("<%dQ" % rank, data, p)) if rank else ()

    @staticmethod
    def _datatype(data):
        cls, ver = data[0] & 0x0F, data[0] >> 4
        bits = data[1] | (data[2] << 8) | (data[3] << 16)
        size = struct.unpack_from("<I", data, 4)[0]
        if cls == 0:
            if bits & 1:
                raise H5Error("big-endian integers")
            return ("int" if bits & 8 else "uint", size)
        if cls == 1:
            if bits & 1:
                raise H5Error("big-endian floats")
            _off, prec, eloc, esize, mloc, msize, bias = struct.unpack_from("<HHBBBBI", data, 8)
            if (size, prec, eloc, esize, mloc, msize, bias) != (8, 64, 52, 11, 0, 52, 1023):
                raise H5Error("not an IEEE f64")
            return ("f64", 8)
        if cls == 3:
            return ("str", size)
        if cls == 9:
            if (bits & 0x0F) != 1:
                raise H5Error("variable-length sequence (not string)")
            if size != 16:
                raise H5Error("vlen element size %d" % size)
            return ("vstr", 16)
        if cls == 6:  # compound: only {r: f64, i: f64} (how EigenHDF5 / h5py store complex numbers)
            n_members = bits & 0xFFFF
            p, members = 8, []
            for _ in range(n_members):
                e = data.index(b"\0", p)
                name = data[p:e].decode()
                p = (p + ((e + 1 - p + 7) & ~7)) if ver < 3 else e + 1
                if ver < 3:
                    off = struct.unpack_from("<I", data, p)[0]
                    p += 4
                    if ver == 1:
                        p += 28  # dimensionality, reserved, permutation, reserved, 4 dimension sizes
                else:
                    nb = max(1, (size.bit_length() + 7) // 8)
                    off = int.from_bytes(data[p:p + nb], "little")
                    p += nb
                mt = File._datatype(data[p:])
                if mt != ("f64", 8):
                    raise H5Error("compound member %s is not f64" % name)
                p += 8 + 12  # an IEEE float datatype message: 8 header bytes + 12 property bytes
                members.append((name, off))
            if size != 16 or members != [("r", 0), ("i", 8)]:
                raise H5Error("compound type %r of %d bytes is not a complex number" % (members, size))
            return ("c128", 16)
        raise H5Error("datatype class %d" % cls)

    def _decode(self, dt, shape, raw):
        kind, size = dt
        count = int(np.prod(shape)) if shape else 1
        if kind == "f64":
            a = np.frombuffer(raw, dtype="<f8", count=count)
        elif kind == "c128":
            a = np.frombuffer(raw, dtype="<c16", count=count)
        elif kind in ("int", "uint"):
            a = np.frombuffer(raw, dtype="<%s%d" % ("i" if kind == "int" else "u", size), count=count)
        elif kind == "str":
            vals = [raw[i * size:(i + 1) * size].split(b"\0")[0].decode() for i in range(count)]
            return vals[0] if not shape else np.array(vals).reshape(shape)
        else:
            vals = []
            for i in range(count):
                ln, addr, idx = struct.unpack_from("<IQI", raw, 16 * i)
                vals.append(self._global_heap_object(addr, idx)[:ln].decode() if ln else "")
            return vals[0] if not shape else np.array(vals).reshape(shape)
        return a[0].item() if not shape else a.reshape(shape).copy()

    def _attribute(self, data):
        ver = data[0]
        if ver == 1:
            ns, dts, dss = struct.unpack_from("<HHH", data, 2)
            p = 8
            pad = lambda n: (n + 7) & ~7
        elif ver in (2, 3):
            ns, dts, dss = struct.unpack_from("<HHH", data, 2)
            p = 8 if ver == 2 else 9
            pad = lambda n: n
        else:
            raise H5Error("attribute version %d" % ver)
        name = data[p:p + ns].split(b"\0")[0].decode()
        p += pad(ns)
        dt = self._datatype(data[p:p + dts])
        p += pad(dts)
        shape = self._dataspace(data[p:p + dss])
        p += pad(dss)
        return name, self._decode(dt, shape, data[p:])

    def _object(self, addr):
        msgs = self._messages(addr)
        types = [m[0] for m in msgs]
        node = Node("group" if 0x0011 in types else "dataset")
        shape = dt = layout = None
        for mtype, data, _flags in msgs:
            if mtype == 0x0011:
                btree, heap = struct.unpack_from("<QQ", data, 0)
                for name, header in self._symbols(btree, heap):
                    node.children[name] = self._object(header)
            elif mtype == 0x0001:
                shape = self._dataspace(data)
            elif mtype == 0x0003:
                dt = self._datatype(data)
            elif mtype == 0x0008:
                layout = data
            elif mtype == 0x000C:
                k, v = self._attribute(data)
                node.attrs[k] = v
        if node.kind == "dataset":
            if dt is None or layout is None or shape is None:
                raise H5Error("dataset at %d lacks dataspace / datatype / layout" % addr)
            if layout[0] != 3:
                raise H5Error("data layout version %d" % layout[0])
            count = int(np.prod(shape)) if shape else 1
            if layout[1] == 1:
                daddr, dsize = struct.unpack_from("<QQ", layout, 2)
                if dsize != count * dt[1]:
                    raise H5Error("layout size %d != %d elements x %d bytes" % (dsize, count, dt[1]))
                raw = self.b[daddr:daddr + dsize] if daddr != UNDEF else b""
                if len(raw) != dsize:
                    raise H5Error("raw data beyond the file")
            elif layout[1] == 0:
                n = struct.unpack_from("<H", layout, 2)[0]
                raw = layout[4:4 + n]
            else:
                raise H5Error("chunked layout not supported")
            node.shape, node.dtype = shape, dt[0]
            node.value = self._decode(dt, shape, raw)
        return node


def tree(path):
    """{path: (kind, dtype, shape)} of every object, for structural comparisons"""
    f = File(path)
    return {p: (n.kind, n.dtype, n.shape) for p, n in f.root.walk()}
