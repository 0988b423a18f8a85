"""Test infrastructure: a minimal tiled / MIP-mapped OpenEXR reader and writer in
numpy (ZIP compression, half / float channels), used to check every level of the
files b2r_tileset_write_exr produces -- OpenCV (which bundles the OpenEXR
library) only decodes level 0.  Pinned against that library by
tests/test_cpu.py::test_exr_mini_against_openexr.  Follows the OpenEXR file
layout document (magic, version flags, attribute list, offset table, tile
chunks) and the library's zip codec (byte interleave + delta predictor + zlib).
Nothing here is used by the product."""
import os
import struct
import zlib

import numpy as np

os.environ.setdefault("OPENCV_IO_ENABLE_OPENEXR", "1")

_NAMES = {1: ["Y"], 2: ["Y", "A"], 3: ["R", "G", "B"], 4: ["R", "G", "B", "A"]}


def zip_block(raw, zlevel=1):
    n = len(raw)
    a = np.frombuffer(raw, np.uint8)
    t = np.concatenate([a[0::2], a[1::2]])  # even bytes, then odd bytes
    d = t.copy()
    d[1:] = t[1:] - t[:-1] + np.uint8(128)   # modulo 256: d = cur - prev + (128 + 256)
    z = zlib.compress(d.tobytes(), zlevel)
    return z if len(z) < n else raw


def unzip_block(z, n):
    if len(z) == n:
        return z  # stored uncompressed
    d = np.frombuffer(zlib.decompress(z), np.uint8)
    assert d.size == n
    t = np.cumsum(np.concatenate([d[:1], d[1:] - np.uint8(128)]), dtype=np.uint8)
    a = np.empty(n, np.uint8)
    a[0::2] = t[:(n + 1) // 2]
    a[1::2] = t[(n + 1) // 2:]
    return a.tobytes()


def level_sizes(w, h):
    out = [(w, h)]
    while w > 1 or h > 1:
        w, h = max(1, w // 2), max(1, h // 2)
        out.append((w, h))
    return out


def write(path, levels, tile=64):
    h, w, c = levels[0].shape
    ptype = {np.dtype(np.float16): 1, np.dtype(np.float32): 2}[levels[0].dtype]
    names = _NAMES[c]
    order = sorted(range(c), key=lambda i: names[i])

    def attr(name, typ, data):
        return name.encode() + b"\0" + typ.encode() + b"\0" + struct.pack("<i", len(data)) + data
    ch = b"".join(names[i].encode() + b"\0" + struct.pack("<iB3xii", ptype, 0, 1, 1)
                  for i in order) + b"\0"
    hdr = (attr("channels", "chlist", ch) + attr("compression", "compression", bytes([3]))
           + attr("dataWindow", "box2i", struct.pack("<4i", 0, 0, w - 1, h - 1))
           + attr("displayWindow", "box2i", struct.pack("<4i", 0, 0, w - 1, h - 1))
           + attr("lineOrder", "lineOrder", bytes([0]))
           + attr("pixelAspectRatio", "float", struct.pack("<f", 1.0))
           + attr("screenWindowCenter", "v2f", struct.pack("<2f", 0, 0))
           + attr("screenWindowWidth", "float", struct.pack("<f", 1.0))
           + attr("tiles", "tiledesc", struct.pack("<IIB", tile, tile, 1)) + b"\0")
    with open(path, "wb") as f:
        f.write(struct.pack("<II", 20000630, 2 | 0x200))
        f.write(hdr)
        n = sum((-(-l.shape[1] // tile)) * (-(-l.shape[0] // tile)) for l in levels)
        table = f.tell()
        f.write(b"\0" * 8 * n)
        offs = []
        for lv, img in enumerate(levels):
            lh, lw, _ = img.shape
            for ty in range(-(-lh // tile)):
                for tx in range(-(-lw // tile)):
                    sub = img[ty * tile:(ty + 1) * tile, tx * tile:(tx + 1) * tile]
                    raw = b"".join(np.ascontiguousarray(sub[y, :, i]).tobytes()
                                   for y in range(sub.shape[0]) for i in order)
                    z = zip_block(raw)
                    offs.append(f.tell())
                    f.write(struct.pack("<5i", tx, ty, lv, lv, len(z)))
                    f.write(z)
        f.seek(table)
        f.write(struct.pack("<%dQ" % len(offs), *offs))


def read(path):
    """-> (list of (h, w, c) level arrays in file channel order re-mapped to
    Y/YA/RGB/RGBA order, dict of attributes {name: (type, bytes)})"""
    b = open(path, "rb").read()
    magic, version = struct.unpack_from("<II", b, 0)
    assert magic == 20000630 and version in (2 | 0x200, 2 | 0x200 | 0x400), hex(version)
    p = 8
    attrs = {}
    while b[p] != 0:
        e = b.index(b"\0", p)
        name = b[p:e].decode()
        p = e + 1
        e = b.index(b"\0", p)
        typ = b[p:e].decode()
        p = e + 1
        n = struct.unpack_from("<i", b, p)[0]
        p += 4
        assert len(name) <= (255 if version & 0x400 else 31), name
        attrs[name] = (typ, b[p:p + n])
        p += n
    p += 1
    for need in ("channels", "compression", "dataWindow", "displayWindow", "lineOrder",
                 "pixelAspectRatio", "screenWindowCenter", "screenWindowWidth", "tiles"):
        assert need in attrs, need
    assert attrs["compression"][1] == bytes([3]) and attrs["lineOrder"][1] == bytes([0])
    chans = []
    q, ch = 0, attrs["channels"][1]
    while ch[q] != 0:
        e = ch.index(b"\0", q)
        nm = ch[q:e].decode()
        pt, plin, xs, ys = struct.unpack_from("<iB3xii", ch, e + 1)
        assert xs == 1 and ys == 1
        chans.append((nm, pt))
        q = e + 1 + 16
    assert [n for n, _ in chans] == sorted(n for n, _ in chans), "channels must be sorted"
    assert len({pt for _, pt in chans}) == 1
    dt = {1: np.float16, 2: np.float32}[chans[0][1]]
    x0, y0, x1, y1 = struct.unpack("<4i", attrs["dataWindow"][1])
    assert (x0, y0) == (0, 0)
    w, h = x1 + 1, y1 + 1
    tw, th, mode = struct.unpack("<IIB", attrs["tiles"][1])
    assert mode == 1  # MIPMAP_LEVELS, ROUND_DOWN
    sizes = level_sizes(w, h)
    c = len(chans)
    want = _NAMES[c]
    ntiles = [(-(-lw // tw)) * (-(-lh // th)) for lw, lh in sizes]
    offs = struct.unpack_from("<%dQ" % sum(ntiles), b, p)
    levels = [np.zeros((lh, lw, c), dt) for lw, lh in sizes]
    k = 0
    for lv, (lw, lh) in enumerate(sizes):
        for ty in range(-(-lh // th)):
            for tx in range(-(-lw // tw)):
                o = offs[k]
                k += 1
                ctx, cty, clx, cly, n = struct.unpack_from("<5i", b, o)
                assert (ctx, cty, clx, cly) == (tx, ty, lv, lv)
                vw, vh = min(tw, lw - tx * tw), min(th, lh - ty * th)
                raw = unzip_block(b[o + 20:o + 20 + n], vw * vh * c * np.dtype(dt).itemsize)
                a = np.frombuffer(raw, dt).reshape(vh, c, vw)
                for j, (nm, _) in enumerate(chans):
                    levels[lv][ty * th:ty * th + vh, tx * tw:tx * tw + vw, want.index(nm)] = a[:, j]
    return levels, attrs


def read_openexr_level0(path, nchannels):
    """Level 0 through the OpenEXR library (OpenCV): (h, w, c) float32, or None."""
    try:
        import cv2
    except ImportError:
        return None
    im = cv2.imread(str(path), cv2.IMREAD_UNCHANGED)
    if im is None:
        return None
    im = im.reshape(im.shape[0], im.shape[1], -1)
    if nchannels == 3:
        im = im[..., ::-1]
    elif nchannels == 4:
        im = im[..., [2, 1, 0, 3]]
    return np.ascontiguousarray(im)
