"""Test infrastructure: a minimal tiled-TIFF reader / writer in numpy, used to
check the files b2r_tileset_write_tiff produces (half samples, which neither
PIL nor OpenCV decode) and pinned itself against libtiff (through OpenCV and
PIL) by tests/test_cpu.py::test_tiff_mini_against_libtiff.  Follows TIFF 6.0 /
BigTIFF and libtiff's fpDiff / fpAcc for Predictor = 3; nothing here is used
by the product."""
import struct
import zlib

import numpy as np

_TYPE = {1: "B", 2: "c", 3: "H", 4: "I", 16: "Q"}


def fp_predict_row(row_bytes, bps, stride):
    """libtiff fpDiff on the bytes of one tile row (little-endian samples)."""
    wc = row_bytes.size // bps
    t = row_bytes.reshape(wc, bps)
    planes = np.concatenate([t[:, bps - 1 - p] for p in range(bps)])  # MSB plane first
    out = planes.copy()
    out[stride:] = planes[stride:] - planes[:-stride]
    return out


def fp_unpredict_rows(rows, bps, stride):
    """libtiff fpAcc on an array (nrows, row_bytes) of predicted bytes."""
    rows = rows.copy()
    for k in range(stride, rows.shape[1]):
        rows[:, k] += rows[:, k - stride]
    wc = rows.shape[1] // bps
    planes = rows.reshape(rows.shape[0], bps, wc)[:, ::-1, :]
    return np.ascontiguousarray(planes.transpose(0, 2, 1)).reshape(rows.shape[0], -1)


def read(path):
    """-> list of dicts, one per directory: {'pixels': (h, w, c) array, 'tags': {tag: value}}"""
    b = open(path, "rb").read()
    assert b[:2] == b"II"
    magic = struct.unpack_from("<H", b, 2)[0]
    big = magic == 43
    assert magic in (42, 43)
    if big:
        assert struct.unpack_from("<HH", b, 4) == (8, 0)
        ifd = struct.unpack_from("<Q", b, 8)[0]
    else:
        ifd = struct.unpack_from("<I", b, 4)[0]
    out = []
    while ifd:
        n = struct.unpack_from("<Q" if big else "<H", b, ifd)[0]
        p = ifd + (8 if big else 2)
        esz, inl = (20, 8) if big else (12, 4)
        tags = {}
        last = -1
        for k in range(n):
            tag, typ = struct.unpack_from("<HH", b, p)
            cnt = struct.unpack_from("<Q" if big else "<I", b, p + 4)[0]
            assert tag > last, "directory entries must be sorted"
            last = tag
            size = struct.calcsize(_TYPE[typ]) * cnt
            voff = p + (12 if big else 8)
            if size > inl:
                voff = struct.unpack_from("<Q" if big else "<I", b, voff)[0]
                assert voff % 2 == 0
            if typ == 2:
                tags[tag] = b[voff:voff + cnt].rstrip(b"\0").decode()
            else:
                tags[tag] = list(struct.unpack_from("<%d%s" % (cnt, _TYPE[typ]), b, voff))
            p += esz
        nxt = struct.unpack_from("<Q" if big else "<I", b, p)[0]
        w, h = tags[256][0], tags[257][0]
        c = tags[277][0]
        bits = tags[258]
        assert len(bits) == c and len(set(bits)) == 1 and tags[339] in ([3] * c, [1] * c)
        assert tags[284] == [1]
        bps = bits[0] // 8
        if tags[339][0] == 3:
            dt = {2: np.float16, 4: np.float32}[bps]
        else:
            dt = {1: np.uint8, 2: np.uint16}[bps]
        tw, th = tags[322][0], tags[323][0]
        ntx, nty = -(-w // tw), -(-h // th)
        offs, cnts = tags[324], tags[325]
        assert len(offs) == len(cnts) == ntx * nty
        full = np.zeros((nty * th, ntx * tw, c), dt)
        pad_zero = True
        for t, (o, m) in enumerate(zip(offs, cnts)):
            raw = b[o:o + m]
            if tags[259] == [8]:
                raw = zlib.decompress(raw)
            else:
                assert tags[259] == [1]
            assert len(raw) == tw * th * c * bps
            rows = np.frombuffer(raw, np.uint8).reshape(th, tw * c * bps)
            if tags.get(317, [1]) == [3]:
                assert tags[339][0] == 3
                rows = fp_unpredict_rows(rows, bps, c)
            px = np.frombuffer(rows.tobytes(), dt).reshape(th, tw, c)
            if tags.get(317, [1]) == [2]:
                assert tags[339][0] == 1
                px = np.cumsum(px, axis=1, dtype=dt)  # libtiff horAcc8 / horAcc16
            ty, tx = divmod(t, ntx)
            full[ty * th:(ty + 1) * th, tx * tw:(tx + 1) * tw] = px
        pad = full.copy()
        pad[:h, :w] = 0
        pad_zero = not pad.view({1: np.uint8, 2: np.uint16, 4: np.uint32}[bps]).any()
        out.append({"pixels": full[:h, :w].copy(), "tags": tags, "big": big,
                    "padding_is_zero": pad_zero})
        ifd = nxt
    return out


def write(path, levels, tile=64, predictor=3, big=False, zlevel=1):
    """levels: list of float32 / float16 arrays (h, w, c) -> tiled TIFF, one
    directory per level (the layout b2r_tileset_write_tiff produces)."""
    f = open(path, "wb")
    if big:
        f.write(struct.pack("<2sHHHQ", b"II", 43, 8, 0, 0))
    else:
        f.write(struct.pack("<2sHI", b"II", 42, 0))
    link = 8 if big else 4
    for img in levels:
        h, w, c = img.shape
        bps = img.dtype.itemsize
        ntx, nty = -(-w // tile), -(-h // tile)
        offs, cnts = [], []
        for ty in range(nty):
            for tx in range(ntx):
                t = np.zeros((tile, tile, c), img.dtype)
                sub = img[ty * tile:(ty + 1) * tile, tx * tile:(tx + 1) * tile]
                t[:sub.shape[0], :sub.shape[1]] = sub
                if predictor == 3:
                    raw = b"".join(fp_predict_row(np.frombuffer(t[r].tobytes(), np.uint8),
                                                  bps, c).tobytes() for r in range(tile))
                elif predictor == 2:  # libtiff horDiff8 / horDiff16
                    d = t.copy()
                    d[:, 1:] = t[:, 1:] - t[:, :-1]
                    raw = d.tobytes()
                else:
                    raw = t.tobytes()
                z = zlib.compress(raw, zlevel)
                if f.tell() & 1:
                    f.write(b"\0")
                offs.append(f.tell())
                cnts.append(len(z))
                f.write(z)
        tags = []

        def T(tag, typ, vals):
            if typ == 2:
                data = vals.encode() + b"\0"
                n = len(data)
            else:
                data = struct.pack("<%d%s" % (len(vals), _TYPE[typ]), *vals)
                n = len(vals)
            tags.append((tag, typ, n, data))
        T(256, 4, [w]); T(257, 4, [h]); T(258, 3, [bps * 8] * c); T(259, 3, [8])
        T(262, 3, [2 if c >= 3 else 1]); T(277, 3, [c]); T(284, 3, [1]); T(305, 2, "tiff_mini")
        if predictor in (2, 3):
            T(317, 3, [predictor])
        T(322, 4, [tile]); T(323, 4, [tile])
        T(324, 16 if big else 4, offs); T(325, 16 if big else 4, cnts)
        ncolor = 3 if c >= 3 else 1
        if c > ncolor:
            T(338, 3, [1] + [0] * (c - ncolor - 1))
        T(339, 3, [3 if img.dtype.kind == "f" else 1] * c)
        T(33302, 2, "Plain Texture"); T(33303, 2, "black,black")
        tags.sort(key=lambda t: t[0])
        inl = 8 if big else 4
        placed = []
        for tag, typ, n, data in tags:
            if len(data) > inl:
                if f.tell() & 1:
                    f.write(b"\0")
                placed.append(f.tell())
                f.write(data)
            else:
                placed.append(None)
        if f.tell() & 1:
            f.write(b"\0")
        ifd = f.tell()
        f.write(struct.pack("<Q" if big else "<H", len(tags)))
        for (tag, typ, n, data), pos in zip(tags, placed):
            f.write(struct.pack("<HH", tag, typ))
            f.write(struct.pack("<Q" if big else "<I", n))
            f.write(data.ljust(inl, b"\0") if pos is None
                    else struct.pack("<Q" if big else "<I", pos))
        nxt = f.tell()
        f.write(struct.pack("<Q" if big else "<I", 0))
        f.seek(link)
        f.write(struct.pack("<Q" if big else "<I", ifd))
        f.seek(0, 2)
        link = nxt
    f.close()


def read_libtiff(path, nchannels):
    """The same file through libtiff (OpenCV; PIL for one float channel): list of
    (h, w, c) arrays, or None when neither decoder is importable."""
    try:
        import cv2
        ok, imgs = cv2.imreadmulti(str(path), flags=cv2.IMREAD_UNCHANGED)
        if ok:
            res = []
            for im in imgs:
                im = im.reshape(im.shape[0], im.shape[1], -1)
                if nchannels == 3:
                    im = im[..., ::-1]          # OpenCV hands back BGR
                elif nchannels == 4:
                    im = im[..., [2, 1, 0, 3]]  # ... and BGRA
                res.append(np.ascontiguousarray(im))
            return res
    except ImportError:
        pass
    if nchannels == 1:
        try:
            from PIL import Image
            im = Image.open(str(path))
            res = []
            for k in range(im.n_frames):
                im.seek(k)
                res.append(np.asarray(im).astype(np.float32)[..., None])
            return res
        except ImportError:
            pass
    return None
