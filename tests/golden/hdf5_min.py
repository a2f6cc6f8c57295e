"""Minimal read-only HDF5 / netCDF-4 reader (no h5py / libhdf5 in this image), enough for the reference's golden test/data/*.nc files:
superblock v2, object headers v2 with continuation chunks, links stored compactly or in a fractal heap (single direct block or
one root indirect block), datasets of fixed-point / IEEE floating-point scalars in compact, contiguous or chunked
(B-tree v1) layout, optional shuffle + deflate filters.  Used only by make_golden.py in the build container; the fixtures it writes are
what travels."""
import struct
import zlib

import numpy as np

UNDEF = 0xFFFFFFFFFFFFFFFF


class H5File:
    def __init__(self, path):
        self.b = open(path, "rb").read()
        b = self.b
        assert b[:8] == b"\x89HDF\r\n\x1a\n", "not an HDF5 file"
        ver = b[8]
        assert ver in (2, 3), f"superblock version {ver} not handled"
        assert b[9] == 8 and b[10] == 8, "only 8-byte offsets / lengths"
        base, ext, eof, root = struct.unpack_from("<QQQQ", b, 12)
        assert base == 0
        self.root = root
        self.datasets = {}
        for name, addr in self._links(root):
            self.datasets[name] = addr

    # ---- object headers -------------------------------------------------------------------------------------------------------
    def _messages(self, addr):
        """[(type, flags, payload bytes)] of the version-2 object header at addr."""
        b = self.b
        assert b[addr:addr + 4] == b"OHDR" and b[addr + 4] == 2, "only version-2 object headers"
        flags = b[addr + 5]
        p = addr + 6
        if flags & 0x20:
            p += 16
        if flags & 0x10:
            p += 4
        nsz = 1 << (flags & 3)
        size0 = int.from_bytes(b[p:p + nsz], "little")
        p += nsz
        track_order = bool(flags & 0x04)
        out = []
        blocks = [(p, size0)]
        while blocks:
            q, n = blocks.pop(0)
            end = q + n
            hdr = 4 + (2 if track_order else 0)
            while q + hdr <= end:
                mtype = b[q]
                msize = struct.unpack_from("<H", b, q + 1)[0]
                mflags = b[q + 3]
                q += hdr
                data = b[q:q + msize]
                q += msize
                if mtype == 0x10:  # continuation: (offset, length) of an OCHK block
                    off, ln = struct.unpack_from("<QQ", data, 0)
                    assert b[off:off + 4] == b"OCHK"
                    blocks.append((off + 4, ln - 8))  # minus signature and checksum
                elif mtype != 0:
                    out.append((mtype, mflags, data))
        return out

    # ---- links ------------------------------------------------------------------------------------------------------------------
    @staticmethod
    def _parse_link(d, p=0):
        """One Link message at d[p:]; returns (name, address | None, next offset)."""
        assert d[p] == 1
        fl = d[p + 1]
        p += 2
        ltype = 0
        if fl & 0x08:
            ltype = d[p]
            p += 1
        if fl & 0x04:
            p += 8
        if fl & 0x10:
            p += 1
        nsz = 1 << (fl & 3)
        ln = int.from_bytes(d[p:p + nsz], "little")
        p += nsz
        name = d[p:p + ln].decode()
        p += ln
        addr = None
        if ltype == 0:
            addr = struct.unpack_from("<Q", d, p)[0]
            p += 8
        elif ltype == 1:  # soft link
            sl = struct.unpack_from("<H", d, p)[0]
            p += 2 + sl
        else:
            raise NotImplementedError("external / user links")
        return name, addr, p

    def _links(self, addr):
        out = []
        for mtype, _, d in self._messages(addr):
            if mtype == 0x06:
                name, a, _ = self._parse_link(d)
                if a is not None:
                    out.append((name, a))
            elif mtype == 0x02:  # Link Info: fractal heap address of densely stored links
                fl = d[1]
                p = 2 + (8 if fl & 1 else 0)
                heap = struct.unpack_from("<Q", d, p)[0]
                if heap != UNDEF:
                    out += self._heap_links(heap)
        return out

    def _heap_links(self, addr):
        b = self.b
        assert b[addr:addr + 4] == b"FRHP" and b[addr + 4] == 0
        p = addr + 5
        heap_id_len, filt_len, flags = struct.unpack_from("<HHB", b, p)
        p += 5
        assert filt_len == 0
        p += 4  # max size of managed objects
        p += 8 * 3  # next huge id, huge b-tree address, free space in managed blocks
        p += 8  # free-space manager address
        p += 8 * 4  # managed space, allocated managed space, iterator offset, number of managed objects
        p += 8 * 4  # huge size / count, tiny size / count
        table_width, start_block, max_direct, max_heap_bits, start_rows, root_addr, cur_rows = struct.unpack_from("<HQQHHQH", b, p)
        off_bytes = (max_heap_bits + 7) // 8
        checksummed = bool(flags & 2)

        def direct(a, size):
            assert b[a:a + 4] == b"FHDB"
            q = a + 5 + 8 + off_bytes + (4 if checksummed else 0)
            links = []
            end = a + size
            while q < end and b[q] == 1:
                try:
                    name, la, q = self._parse_link(b, q)
                except Exception:
                    break
                if la is not None and name:
                    links.append((name, la))
            return links

        if cur_rows == 0:
            return direct(root_addr, start_block)
        # root indirect block: signature, version, heap header address, block offset, then child direct block addresses
        assert b[root_addr:root_addr + 4] == b"FHIB"
        q = root_addr + 5 + 8 + off_bytes
        links = []
        for row in range(cur_rows):
            size = start_block * (1 if row < 2 else 1 << (row - 1))
            for _ in range(table_width):
                a = struct.unpack_from("<Q", b, q)[0]
                q += 8
                if a != UNDEF and size <= max_direct:
                    links += direct(a, size)
        return links

    # ---- datasets -----------------------------------------------------------------------------------------------------------------
    def read(self, name):
        b = self.b
        msgs = self._messages(self.datasets[name])
        shape = dtype = layout = None
        filters = []
        for mtype, _, d in msgs:
            if mtype == 0x01:
                ver, rank, fl = d[0], d[1], d[2]
                p = 8 if ver == 1 else 4
                shape = struct.unpack_from("<" + "Q" * rank, d, p) if rank else ()
            elif mtype == 0x03:
                cls, size = d[0] & 0x0F, struct.unpack_from("<I", d, 4)[0]
                bo = ">" if d[1] & 1 else "<"
                if cls == 0:
                    dtype = np.dtype(bo + ("i" if d[1] & 0x08 else "u") + str(size))
                elif cls == 1:
                    dtype = np.dtype(bo + "f" + str(size))
                else:
                    dtype = np.dtype("V" + str(size))  # strings etc.: returned raw
            elif mtype == 0x08:
                layout = d
            elif mtype == 0x0B:
                ver, nf = d[0], d[1]
                p = 8 if ver == 1 else 2
                for _ in range(nf):
                    fid = struct.unpack_from("<H", d, p)[0]
                    p += 2
                    nlen = 0
                    if ver == 1 or fid >= 256:
                        nlen = struct.unpack_from("<H", d, p)[0]
                        p += 2
                    _, ncd = struct.unpack_from("<HH", d, p)
                    p += 4
                    if nlen:
                        p += (nlen + 7) // 8 * 8 if ver == 1 else nlen
                    p += 4 * ncd
                    if ver == 1 and ncd % 2:
                        p += 4
                    filters.append(fid)
        assert shape is not None and dtype is not None and layout is not None, name
        n = int(np.prod(shape)) if shape else 1
        ver, cls = layout[0], layout[1]
        assert ver == 3, f"layout version {ver}"
        if cls == 0:
            sz = struct.unpack_from("<H", layout, 2)[0]
            return np.frombuffer(layout[4:4 + sz], dtype=dtype, count=n).reshape(shape).copy()
        if cls == 1:
            addr, sz = struct.unpack_from("<QQ", layout, 2)
            if addr == UNDEF:
                return np.zeros(shape, dtype)
            return np.frombuffer(b, dtype=dtype, count=n, offset=addr).reshape(shape).copy()
        assert cls == 2
        nd = layout[2]
        btree = struct.unpack_from("<Q", layout, 3)[0]
        cdims = struct.unpack_from("<" + "I" * nd, layout, 11)
        rank = nd - 1
        assert rank == len(shape) and cdims[-1] == dtype.itemsize
        out = np.zeros(shape, dtype)
        if btree == UNDEF:
            return out
        chunk_shape = cdims[:rank]
        for offs, addr, size in self._chunks(btree, nd):
            raw = b[addr:addr + size]
            for fid in reversed(filters):
                if fid == 1:
                    raw = zlib.decompress(raw)
                elif fid == 2:
                    a = np.frombuffer(raw, np.uint8).reshape(dtype.itemsize, -1)
                    raw = a.T.tobytes()
                elif fid == 3:  # fletcher32: checksum appended
                    raw = raw[:-4]
                else:
                    raise NotImplementedError(f"filter {fid}")
            c = np.frombuffer(raw, dtype=dtype, count=int(np.prod(chunk_shape))).reshape(chunk_shape)
            sl_out, sl_in = [], []
            for o, cs, s in zip(offs, chunk_shape, shape):
                hi = min(o + cs, s)
                sl_out.append(slice(o, hi))
                sl_in.append(slice(0, hi - o))
            if all(s.stop > s.start for s in sl_out):
                out[tuple(sl_out)] = c[tuple(sl_in)]
        return out

    def _chunks(self, addr, nd):
        b = self.b
        assert b[addr:addr + 4] == b"TREE" and b[addr + 4] == 1
        level, used = b[addr + 5], struct.unpack_from("<H", b, addr + 6)[0]
        p = addr + 8 + 16
        keysz = 8 + 8 * nd
        for _ in range(used):
            size, _mask = struct.unpack_from("<II", b, p)
            offs = struct.unpack_from("<" + "Q" * nd, b, p + 8)
            child = struct.unpack_from("<Q", b, p + keysz)[0]
            p += keysz + 8
            if level == 0:
                yield offs[:-1], child, size
            else:
                yield from self._chunks(child, nd)


if __name__ == "__main__":
    import sys

    f = H5File(sys.argv[1])
    for k in f.datasets:
        a = f.read(k)
        print(k, a.shape, a.dtype, a.ravel()[:4] if a.dtype.kind in "fiu" else "")
