"""GPU parity for the maketx pixel pipeline around the MIP chain:
fixNonFinite / check_nan and make_texture (fixnan -> power-of-two resize ->
float top level -> MIP chain with highlight compensation / sharpening),
against the oracle stages chained the way make_texture_impl chains them
(maketexture.cpp:1666-1705, 1791-1877, 811-987)."""
import numpy as np
import pytest

import synth

pytestmark = pytest.mark.gpu


def _sprinkle(a, seed, n=40):
    """Isolated NaN / +-Inf values (no two within a 3x3 window of each other:
    there the reference's in-place scan and any parallel repair agree)."""
    rng = np.random.RandomState(seed)
    h, w, c = a.shape
    used = np.zeros((h, w), bool)
    bad = [np.nan, np.inf, -np.inf]
    k = 0
    while k < n:
        y, x = rng.randint(0, h), rng.randint(0, w)
        if used[max(0, y - 2):y + 3, max(0, x - 2):x + 3].any():
            continue
        used[y, x] = True
        a[y, x, rng.randint(0, c)] = bad[k % 3]
        k += 1
    # corners and edges: clipped neighbourhoods
    for (y, x) in [(0, 0), (0, w - 1), (h - 1, 0), (h - 1, w - 1)]:
        if not used[max(0, y - 2):y + 3, max(0, x - 2):x + 3].any():
            a[y, x, 0] = np.nan
            used[y, x] = True
            k += 1
    return k


@pytest.mark.parametrize("dtype", [np.float32, np.float16])
@pytest.mark.parametrize("mode", ["box3", "black", "none"])
def test_fix_nonfinite(oiio, oracle, dtype, mode):
    src = synth.image(61, 47, 3, dtype, seed=5, hdr=True)
    n = _sprinkle(src, 7)
    ref = np.zeros_like(src)
    nref = oracle.fix_nonfinite(oracle.Img(ref), oracle.Img(src), mode)
    assert nref == n
    IBA = oiio.ImageBufAlgo
    m = {"box3": oiio.NONFINITE_BOX3, "black": oiio.NONFINITE_BLACK,
         "none": oiio.NONFINITE_NONE}[mode]
    out = IBA.fixNonFinite(oiio.ImageBuf(src), m)
    assert not out.has_error, out.geterror()
    assert IBA.last_pixels_fixed == n
    assert np.array_equal(out.pixels.view(np.uint8), ref.view(np.uint8))
    if mode != "none":
        assert np.isfinite(out.pixels.astype(np.float32)).all()


def test_fix_nonfinite_in_place_device_roi_and_error_mode(oiio, oracle):
    import torch
    src = synth.image(64, 40, 4, np.float32, seed=9)
    _sprinkle(src, 11, 25)
    ref = src.copy()
    roi = (4, 60, 3, 37)
    nref = oracle.fix_nonfinite(oracle.Img(ref), oracle.Img(ref), "box3", roi=roi,
                                chrange=(0, 3))
    t = torch.from_numpy(src).cuda()
    B = oiio.ImageBuf(t)
    IBA = oiio.ImageBufAlgo
    assert IBA.fixNonFinite(B, B, oiio.NONFINITE_BOX3,
                            roi=oiio.ROI(roi[0], roi[1], roi[2], roi[3], 0, 1, 0, 3))
    assert IBA.last_pixels_fixed == nref
    got = t.cpu().numpy()
    assert np.array_equal(got.view(np.uint32), ref.view(np.uint32))
    # NONFINITE_ERROR: counts and fails
    bad = IBA.fixNonFinite(oiio.ImageBuf(src), oiio.NONFINITE_ERROR)
    assert bad.geterror() == "Nonfinite pixel values found"
    # integer images: a plain copy
    u8 = synth.image(30, 20, 3, np.uint8, seed=3)
    out = IBA.fixNonFinite(oiio.ImageBuf(u8), oiio.NONFINITE_BOX3)
    assert np.array_equal(out.pixels, u8) and IBA.last_pixels_fixed == 0


def _oracle_pipeline(oracle, src, cfg):
    """make_texture_impl's pixel stages with the oracle."""
    a = src
    if cfg.get("maketx:fixnan", "none") != "none" and a.dtype in (np.float32, np.float16):
        b = np.zeros_like(a)
        oracle.fix_nonfinite(oracle.Img(b), oracle.Img(a), cfg["maketx:fixnan"])
        a = b
    h, w, c = a.shape
    f = np.zeros((h, w, c), np.float32)
    oracle.resample(oracle.Img(f), oracle.Img(a), False)       # copy_pixels to float
    fn = cfg.get("maketx:filtername", "box")
    if cfg.get("maketx:resize"):
        p2 = lambda v: 1 << (int(v) - 1).bit_length()
        top = np.zeros((p2(h), p2(w), c), np.float32)
        if fn in ("box", "triangle"):
            oracle.resize_block(oracle.Img(top), oracle.Img(f))
        else:
            oracle.resize(oracle.Img(top), oracle.Img(f), fn)
    else:
        top = f
    sharpen_first = True
    sfilt = "gaussian"
    if fn.startswith("post-"):
        sharpen_first, fn = False, fn[5:]
    if fn.startswith("unsharp-"):
        sfilt, fn = fn[8:], "lanczos3"
    levels = oracle.mip_chain(top, fn, highlightcomp=bool(cfg.get("maketx:highlightcomp")),
                              alpha_channel=3 if c == 4 else -1,
                              sharpen=cfg.get("maketx:sharpen", 0.0),
                              sharpen_first=sharpen_first, sharpenfilt=sfilt,
                              clamp_half=cfg.get("format") == "half")
    return [top] + levels


@pytest.mark.parametrize("cfg", [
    {},                                                             # maketx defaults: box
    {"maketx:filtername": "lanczos3", "maketx:resize": 1},
    {"maketx:filtername": "box", "maketx:resize": 1},
    {"maketx:filtername": "lanczos3", "maketx:highlightcomp": 1, "maketx:fixnan": "box3"},
    {"maketx:filtername": "post-unsharp-gaussian", "maketx:sharpen": 0.5},
    {"maketx:filtername": "blackman-harris", "maketx:sharpen": 0.4, "format": "half"},
])
@pytest.mark.parametrize("dtype,nch", [(np.float32, 4), (np.uint8, 3)])
def test_make_texture_matches_oracle_pipeline(oiio, oracle, cfg, dtype, nch):
    src = synth.image(100, 60, nch, dtype, seed=131)
    if dtype == np.float32:
        src = (src * np.float32(3)).astype(np.float32)
        if cfg.get("maketx:fixnan"):
            _sprinkle(src, 17, 12)
    ref = _oracle_pipeline(oracle, src, cfg)
    T = oiio.make_texture(oiio.MakeTxTexture, oiio.ImageBuf(src.copy()), cfg)
    assert T.ok, T.error
    assert len(T.levels) == len(ref)
    for i, (g, r) in enumerate(zip(T.levels, ref)):
        gp = g.get_pixels()
        assert gp.shape == r.shape, i
        scale = max(1.0, float(np.abs(r).max()))
        tol = (1e-4 if cfg.get("maketx:highlightcomp") else 2e-5) * max(1, i) * scale
        assert np.abs(gp - r).max() <= tol, "level %d" % i


def test_make_texture_checknan_and_nomipmap(oiio):
    src = synth.image(32, 32, 3, np.float32, seed=1)
    T = oiio.make_texture(oiio.MakeTxTexture, oiio.ImageBuf(src), {"maketx:nomipmap": 1})
    assert T.ok and len(T.levels) == 1
    src[5, 6, 1] = np.inf
    src[9, 9, 0] = np.nan
    T = oiio.make_texture(oiio.MakeTxTexture, oiio.ImageBuf(src), {"maketx:checknan": 1})
    assert T.error == "maketx ERROR: Nan/Inf at 2 pixels"
    T = oiio.make_texture(oiio.MakeTxTexture, oiio.ImageBuf(src),
                          {"maketx:checknan": 1, "maketx:fixnan": "black"})
    assert T.ok and T.pixels_fixed == 2


# ------------------------------------------------------ fillholes_pushpull ----
def test_fillholes_golden(oiio, oracle):
    """testsuite/oiiotool: src/hole.tif --fillholes -> ref/tahoe-filled.tif (the
    oracle composition reproduces it bit for bit, test_cpu)."""
    import os
    G = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden",
                             "fillholes_ref.npz"))
    out = oiio.ImageBufAlgo.fillholes_pushpull(oiio.ImageBuf(G["hole"].copy()))
    assert not out.has_error, out.geterror()
    d = np.abs(out.pixels.astype(int) - G["tahoe_filled"].astype(int))
    assert d.max() <= 1 and (d != 0).mean() < 1e-3


@pytest.mark.parametrize("dtype", [np.float32, np.uint16])
@pytest.mark.parametrize("size", [(97, 61), (64, 64), (33, 1)])
def test_fillholes_matches_oracle(oiio, oracle, dtype, size):
    w, h = size
    src = synth.image(w, h, 4, np.float32, seed=141)
    hole = synth.image(w, h, 1, np.float32, seed=142)[..., 0] < 0.35
    src[..., 3] = 1.0
    src[hole] = 0.0                                   # premultiplied holes
    if dtype != np.float32:
        src = oracle.quantize(src, dtype)
    ref = oracle.fillholes_pushpull(src, 3)
    B = oiio.ImageBuf(src.copy())
    B.spec().x, B.spec().y = 7, -3                    # data origin off the display origin
    out = oiio.ImageBufAlgo.fillholes_pushpull(B)
    assert not out.has_error, out.geterror()
    if dtype == np.float32:
        assert np.abs(out.pixels - ref).max() <= 5e-5
        assert (out.pixels[..., 3] > 0.99).all() or hole.all()
    else:
        d = np.abs(out.pixels.astype(int) - ref.astype(int))
        assert d.max() <= 3 and (d > 1).mean() < 1e-3


def test_fillholes_errors(oiio):
    rgb = oiio.ImageBuf(np.zeros((8, 8, 3), np.float32))
    assert oiio.ImageBufAlgo.fillholes_pushpull(rgb).geterror() == \
        "images must have alpha channels"


# ------------------------------------------------------------- pixel stats ----
@pytest.mark.parametrize("dtype,nch", [(np.float32, 4), (np.float32, 3), (np.float16, 4),
                                       (np.uint8, 3), (np.uint16, 1), (np.float32, 6)])
def test_pixel_stats(oiio, oracle, dtype, nch):
    src = synth.image(333, 257, nch, dtype, seed=151, hdr=(dtype == np.float32))
    if dtype in (np.float32, np.float16):
        src[5, 7, 0] = np.nan
        src[100, 200, nch - 1] = np.inf
        src[101, 200, nch - 1] = -np.inf
    ref = oracle.pixel_stats(src)
    st = oiio.ImageBufAlgo.computePixelStats(oiio.ImageBuf(src))
    assert len(st.min) == nch
    for c in range(nch):
        r = ref[c]
        assert st.min[c] == r["min"] and st.max[c] == r["max"], c
        assert (st.nancount[c], st.infcount[c], st.finitecount[c]) == \
            (r["nancount"], r["infcount"], r["finitecount"]), c
        assert abs(st.sum[c] - r["sum"]) <= 1e-11 * max(1.0, abs(r["sum"])), c
        assert abs(st.sum2[c] - r["sum2"]) <= 1e-11 * max(1.0, abs(r["sum2"])), c
        assert abs(st.avg[c] - r["avg"]) <= 1e-6 * max(1.0, abs(r["avg"])), c
        assert abs(st.stddev[c] - r["stddev"]) <= 1e-5 * max(1.0, abs(r["stddev"])), c
    assert st.monochrome == oracle.is_monochrome(src)


def test_pixel_stats_reference_kat_monochrome_constant_device(oiio, oracle):
    """The reference's own known answer (imagebufalgo_test.cpp:1101-1123), the
    monochrome / constant-colour answers, a device-resident image, a ROI."""
    import torch
    IBA = oiio.ImageBufAlgo
    img = np.zeros((2, 2, 3), np.float32)
    img[:, 1, :] = 1.0
    st = IBA.computePixelStats(oiio.ImageBuf(img))
    for c in range(3):
        assert (st.min[c], st.max[c], st.avg[c], st.stddev[c]) == (0.0, 1.0, 0.5, 0.5)
        assert (st.nancount[c], st.infcount[c], st.finitecount[c]) == (0, 0, 4)
    assert st.monochrome and IBA.isConstantColor(oiio.ImageBuf(img)) is None
    const = np.full((40, 50, 4), 0.25, np.float32)
    assert IBA.isConstantColor(oiio.ImageBuf(const)) == [0.25] * 4
    grey = synth.image(64, 48, 1, np.uint8, seed=3).repeat(3, axis=2)
    assert IBA.isMonochrome(oiio.ImageBuf(grey))
    grey[47, 63, 2] ^= 1
    assert not IBA.isMonochrome(oiio.ImageBuf(grey))
    f = synth.image(64, 48, 3, np.float32, seed=4)
    f[..., 1] = f[..., 0] + np.float32(0.001)
    f[..., 2] = f[..., 0]
    assert not IBA.isMonochrome(oiio.ImageBuf(f))
    assert IBA.isMonochrome(oiio.ImageBuf(f), 0.01) == oracle.is_monochrome(f, 0.01) == True
    big = synth.image(1024, 512, 4, np.float32, seed=5)
    t = torch.from_numpy(big).cuda()
    st = IBA.computePixelStats(oiio.ImageBuf(t), roi=oiio.ROI(10, 700, 20, 500, 0, 1, 1, 3))
    ref = oracle.pixel_stats(big[20:500, 10:700, 1:3].copy())
    for k, c in enumerate((1, 2)):
        assert st.min[c] == ref[k]["min"] and st.finitecount[c] == ref[k]["finitecount"]
        assert abs(st.sum[c] - ref[k]["sum"]) <= 1e-11 * abs(ref[k]["sum"])
    assert st.finitecount[0] == 0 and st.min[0] == 0.0      # channel outside the ROI


def test_make_texture_reports_statistics(oiio, oracle):
    src = synth.image(64, 32, 3, np.float32, seed=6)
    T = oiio.make_texture(oiio.MakeTxTexture, oiio.ImageBuf(src), {"maketx:nomipmap": 1})
    ref = oracle.pixel_stats(src)
    assert T.ok and T.constant_color is None and T.monochrome is False
    assert all(abs(a - r["avg"]) < 1e-6 for a, r in zip(T.average_color, ref))


@pytest.mark.parametrize("fmt", ["float", "half"])
def test_make_texture_writes_a_tiff_texture(oiio, tmp_path, fmt):
    """make_texture with an output file name: the levels it returns are the
    levels in the file (tiled TIFF, one directory per level, Predictor 3), and the
    ImageDescription carries the SHA-1 of the top level and the average colour
    (maketexture.cpp:1928-1975)."""
    import hashlib
    import tiff_mini
    src = synth.image(300, 200, 4, np.float32, seed=77)
    path = tmp_path / "out.tx.tif"
    T = oiio.make_texture(oiio.MakeTxTexture, oiio.ImageBuf(src),
                          {"maketx:filtername": "lanczos3", "format": fmt},
                          outputfilename=path)
    assert T.ok, T.error
    dirs = tiff_mini.read(path)
    assert len(dirs) == len(T.levels)
    for lv, d in enumerate(dirs):
        want = np.asarray(T.levels[lv].pixels)
        if fmt == "half":
            want = want.astype(np.float16)
        assert np.array_equal(d["pixels"], want), lv
        assert d["tags"][317] == [3] and d["tags"][322] == [64]
    desc = dirs[0]["tags"][270]
    assert "oiio:SHA-1=" + oiio.ImageBufAlgo.computePixelHashSHA1(
        T.levels[0], "lanczos3 ", blocksize=256) in desc
    assert "oiio:AverageColor=" in desc
    avg = [float(x) for x in desc.split("oiio:AverageColor=")[1].split()[0].split(",")]
    assert np.allclose(avg, src.reshape(-1, 4).mean(0), atol=1e-5)
    if fmt == "float":
        theirs = tiff_mini.read_libtiff(path, 4)
        if theirs is not None:
            assert all(np.array_equal(a, np.asarray(l.pixels)) for a, l in zip(theirs, T.levels))
    bad = oiio.make_texture(oiio.MakeTxTexture, oiio.ImageBuf(src), {"format": "double"},
                            outputfilename=path)
    assert not bad.ok and "tiles are float, half, uint16 or uint8" in bad.error


@pytest.mark.parametrize("half", [False, True])
@pytest.mark.parametrize("nch,shape,tile", [(4, (300, 500), 64), (3, (512, 512), 64),
                                            (1, (97, 130), 32), (2, (70, 41), 16)])
def test_write_exr_texture(oiio, tmp_path, nch, shape, tile, half):
    """The tile set as a tiled, MIP-mapped, zip-compressed OpenEXR file: the
    OpenEXR library (through OpenCV) decodes level 0 bit for bit; every level
    goes through the numpy reader test_cpu pins against that library."""
    import exr_mini
    h, w = shape
    src = synth.image(w, h, nch, np.float32, seed=980 + nch, hdr=True)
    ch = oiio.MipChain(src, filtername="lanczos3", clamp_half=half)
    ts = oiio.TileSet(ch, tile=tile, half=half, compression=False)
    path = tmp_path / "tex.exr"
    assert ts.write_exr(path, attributes={"textureformat": "Plain Texture",
                                          "oiio:SHA-1": "0123ABCD"})
    levels, attrs = exr_mini.read(path)
    assert len(levels) == ch.nlevels
    for lv, got in enumerate(levels):
        px = ch.download(lv)
        want = px.astype(np.float16) if half else px
        assert got.dtype == want.dtype and np.array_equal(got, want), "level %d" % lv
    assert attrs["textureformat"] == ("string", b"Plain Texture")
    assert attrs["oiio:SHA-1"] == ("string", b"0123ABCD")
    if nch != 2:
        l0 = exr_mini.read_openexr_level0(path, nch)
        if l0 is not None:
            want0 = ch.download(0)
            want0 = want0.astype(np.float16).astype(np.float32) if half else want0
            assert np.array_equal(l0, want0)
    with pytest.raises(oiio.B2RError, match="stored"):
        oiio.TileSet(ch, tile=tile, half=half).write_exr(path)
    with pytest.raises(oiio.B2RError, match="every level"):
        oiio.TileSet(ch, tile=tile, compression=False, first_level=1).write_exr(path)


def test_make_texture_writes_an_exr_texture(oiio, tmp_path):
    import exr_mini
    src = synth.image(300, 200, 4, np.float32, seed=79)
    path = tmp_path / "out.exr"
    T = oiio.make_texture(oiio.MakeTxTexture, oiio.ImageBuf(src),
                          {"maketx:filtername": "lanczos3", "format": "half"},
                          outputfilename=path)
    assert T.ok, T.error
    levels, attrs = exr_mini.read(path)
    assert len(levels) == len(T.levels)
    for got, lvl in zip(levels, T.levels):
        assert np.array_equal(got, np.asarray(lvl.pixels).astype(np.float16))
    assert attrs["oiio:SHA-1"][1].decode() == T.sha1
    assert attrs["wrapmodes"] == ("string", b"black,black")
    assert "oiio:AverageColor" in attrs
    bad = oiio.make_texture(oiio.MakeTxTexture, oiio.ImageBuf(src), {"format": "uint8"},
                            outputfilename=path)
    assert not bad.ok and "OpenEXR textures are float or half" in bad.error


def _convert_type_from_float(px, dt):
    """convert_type<float,D> (fmath.h:753-764, 1009-1031) in numpy float32 steps:
    v * max, + 0.5, clamp, truncate."""
    mx = np.float32(np.iinfo(dt).max)
    s = px.astype(np.float32) * mx
    s = s + np.float32(0.5)
    return np.clip(s, np.float32(0), mx).astype(dt)


@pytest.mark.parametrize("dt", [np.uint8, np.uint16])
def test_make_texture_writes_an_integer_tiff_texture(oiio, tmp_path, dt):
    """maketx of a uint8 / uint16 picture: float levels, written through
    convert_type<float,D> as 64x64 deflate tiles with the horizontal predictor
    (tiffoutput.cpp:691-694).  libtiff reads every level back equal to the
    quantised levels make_texture returned."""
    import tiff_mini
    src = synth.image(260, 190, 3, dt, seed=78)
    path = tmp_path / "out8.tif"
    T = oiio.make_texture(oiio.MakeTxTexture, oiio.ImageBuf(src),
                          {"maketx:filtername": "lanczos3"}, outputfilename=path)
    assert T.ok, T.error
    assert T.format == np.dtype(dt).name
    dirs = tiff_mini.read(path)
    assert len(dirs) == len(T.levels)
    want = [_convert_type_from_float(np.asarray(l.pixels), dt) for l in T.levels]
    assert np.array_equal(want[0], src)     # level 0 survives the float round trip
    for lv, d in enumerate(dirs):
        assert d["pixels"].dtype == dt and np.array_equal(d["pixels"], want[lv]), lv
        assert d["tags"][317] == [2] and d["tags"][339] == [1, 1, 1]
        assert d["tags"][258] == [8 * np.dtype(dt).itemsize] * 3
    theirs = tiff_mini.read_libtiff(path, 3)
    if theirs is not None:
        assert len(theirs) == len(want)
        assert all(np.array_equal(a, b) for a, b in zip(theirs, want))


@pytest.mark.parametrize("dt", [np.uint8, np.uint16])
@pytest.mark.parametrize("predictor", [0, 2])
@pytest.mark.parametrize("nch,shape,tile", [(4, (300, 500), 64), (3, (130, 97), 32), (1, (70, 200), 16)])
def test_encode_tiles_integer(oiio, tmp_path, nch, shape, tile, predictor, dt):
    """uint8 / uint16 tiles: every tile decodes to convert_type<float,D> of the
    level (values outside [0, 1] included: lanczos3 overshoots), zero padded; the
    device-to-device entry gives the same bytes; the TIFF goes through libtiff."""
    import zlib
    import torch
    import tiff_mini
    h, w = shape
    src = synth.image(w, h, nch, np.float32, seed=970 + nch) * 1.4 - 0.2
    ch = oiio.MipChain(src, filtername="lanczos3")
    name = np.dtype(dt).name
    ts = oiio.TileSet(ch, tile=tile, out_type=name, predictor=predictor, nthreads=4)
    for lv in range(ch.nlevels):
        q = _convert_type_from_float(ch.download(lv), dt)
        lw, lh, ntx, nty = ts.level_info(lv)
        ref = np.zeros((nty * tile, ntx * tile, nch), dt)
        ref[:lh, :lw] = q
        for ty in range(nty):
            for tx in range(ntx):
                assert np.array_equal(ts.tile_pixels(lv, tx, ty),
                                      ref[ty * tile:(ty + 1) * tile, tx * tile:(tx + 1) * tile]), \
                    (lv, tx, ty)
    dev = oiio.tile_relayout(torch.from_numpy(src).cuda(), tile=tile, out_type=name,
                             predictor=predictor).cpu().numpy()
    _, _, ntx, nty = ts.level_info(0)
    for ty in range(nty):
        for tx in range(ntx):
            assert np.array_equal(dev[ty, tx], np.frombuffer(
                zlib.decompress(ts.tile_bytes(0, tx, ty)), np.uint8))
    path = tmp_path / "i.tif"
    ts.write_tiff(path)
    dirs = tiff_mini.read(path)
    for lv, d in enumerate(dirs):
        assert np.array_equal(d["pixels"], _convert_type_from_float(ch.download(lv), dt))
        assert d["padding_is_zero"]
    theirs = tiff_mini.read_libtiff(path, nch)
    if theirs is not None:
        assert all(np.array_equal(a, d["pixels"]) for a, d in zip(theirs, dirs))
    with pytest.raises(oiio.B2RError, match="predictor 2 goes with"):
        oiio.TileSet(ch, tile=tile, out_type=name, predictor=3)


# ---- the chain cut into row bands (b2r_mipchain_create_banded) ------------
@pytest.mark.gpu
@pytest.mark.parametrize("filt", ["box", "lanczos3", "triangle"])
@pytest.mark.parametrize("shape,nbands", [((1024, 768), 2), ((1100, 900), 3),
                                          ((2048, 2048), 4), ((640, 1500), 2)])
def test_banded_chain_equals_single_device_chain(oiio, filt, shape, nbands):
    """Row bands + pushed halo rows + gathered tail: every level is
    bit-identical to the single-device chain.  The bands all live on device 0
    here (the band / halo / gather logic is the same; real peer copies are
    exercised by tools/banded_gpu.py on a multi-GPU box)."""
    h, w = shape
    src = synth.image(w, h, 4, np.float32, seed=500 + nbands)
    one = oiio.MipChain(src, filtername=filt)
    many = oiio.MipChainBanded(src, [0] * nbands, filtername=filt)
    assert many.nlevels == one.nlevels
    for lv in range(one.nlevels):
        a, b = one.download(lv), many.download(lv)
        assert a.shape == b.shape
        assert np.array_equal(a, b), "level %d differs (max %g)" % (lv, np.abs(a - b).max())
    if nbands > 1 and h // nbands >= 128:
        assert many.ms("halo_bytes") > 0


# ---- tiled, deflated level output (b2r_mipchain_encode_tiles) --------------
@pytest.mark.parametrize("half", [False, True])
@pytest.mark.parametrize("shape,tile", [((300, 500), 64), ((1024, 1024), 64), ((97, 130), 32)])
def test_encode_tiles_roundtrip(oiio, shape, tile, half):
    """Every level, re-laid into tiles on the device, downloaded and deflated
    on the host, inflates back to exactly the level's pixels (float) or their
    half conversion; edge tiles are zero padded."""
    h, w = shape
    src = synth.image(w, h, 3, np.float32, seed=900 + tile)
    ch = oiio.MipChain(src, filtername="lanczos3")
    ts = oiio.TileSet(ch, tile=tile, half=half, zlevel=1, nthreads=4)
    assert ts.nlevels == ch.nlevels
    for lv in range(ch.nlevels):
        px = ch.download(lv)
        lw, lh, ntx, nty = ts.level_info(lv)
        assert (lh, lw) == px.shape[:2]
        assert ntx == -(-lw // tile) and nty == -(-lh // tile)
        ref = np.zeros((nty * tile, ntx * tile, 3), np.float16 if half else np.float32)
        ref[:lh, :lw] = px.astype(np.float16) if half else px
        for ty in range(nty):
            for tx in range(ntx):
                t = ts.tile_pixels(lv, tx, ty)
                assert np.array_equal(t, ref[ty * tile:(ty + 1) * tile, tx * tile:(tx + 1) * tile])
    assert ts.stat("stored_bytes") > 0 and ts.stat("raw_bytes") >= ts.stat("stored_bytes") * 0.5


@pytest.mark.gpu
@pytest.mark.parametrize("half", [False, True])
@pytest.mark.parametrize("nch,shape,tile", [(3, (300, 500), 64), (4, (257, 130), 64),
                                            (1, (200, 333), 32), (2, (97, 130), 16)])
def test_encode_tiles_fp_predictor(oiio, nch, shape, tile, half):
    """predictor = 3: the re-layout kernel applies TIFF's floating-point
    predictor per tile row (tiffoutput.cpp:683-690; libtiff fpDiff).  The stored
    bytes equal the numpy restatement of fpDiff applied to the plain tiles, byte
    for byte, and decode back to the level."""
    import zlib
    import tiff_mini
    h, w = shape
    src = synth.image(w, h, nch, np.float32, seed=930 + tile + nch)
    ch = oiio.MipChain(src, filtername="lanczos3")
    plain = oiio.TileSet(ch, tile=tile, half=half, zlevel=1, nthreads=4)
    pred = oiio.TileSet(ch, tile=tile, half=half, zlevel=1, nthreads=4, predictor=3)
    bps = 2 if half else 4
    for lv in range(ch.nlevels):
        _, _, ntx, nty = plain.level_info(lv)
        for ty in range(nty):
            for tx in range(ntx):
                raw = np.frombuffer(zlib.decompress(plain.tile_bytes(lv, tx, ty)), np.uint8)
                want = np.concatenate([tiff_mini.fp_predict_row(r, bps, nch)
                                       for r in raw.reshape(tile, -1)])
                got = np.frombuffer(zlib.decompress(pred.tile_bytes(lv, tx, ty)), np.uint8)
                assert np.array_equal(got, want), (lv, tx, ty)
                assert np.array_equal(pred.tile_pixels(lv, tx, ty).view(np.uint8),
                                      plain.tile_pixels(lv, tx, ty).view(np.uint8))
    with pytest.raises(oiio.B2RError, match="predictor"):
        oiio.TileSet(ch, tile=tile, predictor=2)


@pytest.mark.gpu
@pytest.mark.parametrize("half", [False, True])
@pytest.mark.parametrize("predictor", [0, 3])
def test_tile_relayout_on_device(oiio, half, predictor):
    """b2r_tile_relayout (device in, device out) produces the bytes the tile
    set stores before deflate, for every tile, ragged edges included."""
    import torch
    import zlib
    src = synth.image(333, 210, 4, np.float32, seed=990)
    dev = torch.from_numpy(src).cuda()
    tiles = oiio.tile_relayout(dev, tile=64, half=half, predictor=predictor).cpu().numpy()
    ch = oiio.MipChain(src, filtername="box")
    ts = oiio.TileSet(ch, tile=64, half=half, predictor=predictor)
    _, _, ntx, nty = ts.level_info(0)
    assert tiles.shape[:2] == (nty, ntx)
    for ty in range(nty):
        for tx in range(ntx):
            want = np.frombuffer(zlib.decompress(ts.tile_bytes(0, tx, ty)), np.uint8)
            assert np.array_equal(tiles[ty, tx], want), (tx, ty)
    with pytest.raises(oiio.B2RError, match="contiguous float32 device image"):
        oiio.tile_relayout(dev[:, ::2], tile=64)
    with pytest.raises(oiio.B2RError, match="multiple of 4 samples"):
        oiio.tile_relayout(dev[:, :, :3].contiguous(), tile=2, predictor=3)


@pytest.mark.gpu
@pytest.mark.parametrize("half,predictor,big", [(False, 3, False), (False, 0, True),
                                                (True, 3, False), (True, 0, False)])
@pytest.mark.parametrize("nch,shape", [(4, (300, 500)), (3, (512, 512)), (1, (97, 130))])
def test_write_tiff_texture(oiio, tmp_path, nch, shape, half, predictor, big):
    """maketx's output file (maketexture.cpp:957-974 + the TIFF plugin): chain ->
    tiles -> a tiled TIFF with one directory per MIP level.  libtiff (OpenCV /
    PIL) reads every float level back bit for bit; half files go through the
    numpy reader that test_cpu pins against libtiff."""
    import tiff_mini
    h, w = shape
    src = synth.image(w, h, nch, np.float32, seed=960 + nch)
    ch = oiio.MipChain(src, filtername="lanczos3")
    ts = oiio.TileSet(ch, tile=64, half=half, zlevel=1, nthreads=4, predictor=predictor)
    path = tmp_path / "tex.tif"
    assert ts.write_tiff(path, image_description="SHA-1=0123", bigtiff=big)
    dirs = tiff_mini.read(path)
    assert len(dirs) == ch.nlevels
    for lv, d in enumerate(dirs):
        px = ch.download(lv)
        want = px.astype(np.float16) if half else px
        assert d["pixels"].shape == want.shape
        assert np.array_equal(d["pixels"], want), "level %d" % lv
        assert d["padding_is_zero"] and d["big"] == big
        t = d["tags"]
        assert t[322] == [64] and t[323] == [64] and t[259] == [8]
        assert t.get(317, [1]) == [predictor or 1]
        assert t[270] == "SHA-1=0123" and t[305] == "openimageio_b200"
        assert t[33302] == "Plain Texture" and t[33303] == "black,black"
        assert t[262] == [2 if nch >= 3 else 1]
        assert (338 in t) == (nch in (2, 4))
    if not half:
        theirs = tiff_mini.read_libtiff(path, nch)
        if theirs is not None:
            assert len(theirs) == ch.nlevels
            for lv, im in enumerate(theirs):
                assert np.array_equal(im, ch.download(lv)), "libtiff level %d" % lv
    # a tile set that skips levels, or tiles TIFF cannot hold, is refused
    with pytest.raises(oiio.B2RError, match="every level"):
        oiio.TileSet(ch, tile=64, first_level=1).write_tiff(path)
    with pytest.raises(oiio.B2RError, match="multiples of 16"):
        oiio.TileSet(ch, tile=40).write_tiff(path)
    with pytest.raises(oiio.B2RError, match="could not open"):
        ts.write_tiff(tmp_path / "no_such_dir" / "x.tif")
    with pytest.raises(oiio.B2RError, match="compressed tiles only"):
        oiio.TileSet(ch, tile=64, compression=False, predictor=3).write_tiff(path)
    # stored (uncompressed) tiles: Compression = 1
    oiio.TileSet(ch, tile=64, half=half, compression=False).write_tiff(path)
    raw = tiff_mini.read(path)
    assert raw[0]["tags"][259] == [1]
    assert all(np.array_equal(a["pixels"], b["pixels"]) for a, b in zip(raw, dirs))


# ---- the chain as one pipeline: upload | level 1 bands | downloads ----------
@pytest.mark.gpu
@pytest.mark.parametrize("filt", ["box", "lanczos3"])
@pytest.mark.parametrize("shape", [(1024, 1536), (700, 1031)])
@pytest.mark.parametrize("pinned", [False, True])
def test_streamed_chain_equals_chain(oiio, filt, shape, pinned):
    """b2r_mipchain_create_streamed delivers every level to host memory while the
    upload of level 0 is still running; each level must be bit-identical to the
    plain chain's (same kernels on the same rows), for even sizes (box: 2x2
    average) and odd ones (box: bilinear at NDC centres)."""
    import torch
    h, w = shape
    lvl0 = synth.image(w, h, 4, np.float32, seed=610)
    host = torch.from_numpy(lvl0).pin_memory() if pinned else lvl0
    ref = oiio.MipChain(lvl0, filt)
    chain, levels = oiio.MipChain.streamed(oiio.ImageBuf(host), filt)
    assert chain.nlevels == ref.nlevels == len(levels) + 1
    for k, got in enumerate(levels, start=1):
        want = ref.download(k)
        assert got.shape == want.shape
        assert np.array_equal(got, want), "level %d differs (max %g)" % (
            k, float(np.abs(got - want).max()))
    # the chain stays resident: its device levels are the same pixels
    assert np.array_equal(chain.download(2), levels[1])


@pytest.mark.gpu
def test_streamed_chain_falls_back(oiio):
    """What the pipeline does not cover (here: a level 0 under 256 rows) takes the
    plain chain followed by the downloads -- same results."""
    lvl0 = synth.image(300, 120, 3, np.float32, seed=611)
    ref = oiio.MipChain(lvl0, "lanczos3")
    chain, levels = oiio.MipChain.streamed(lvl0, "lanczos3")
    for k, got in enumerate(levels, start=1):
        assert np.array_equal(got, ref.download(k))


@pytest.mark.gpu
def test_trim_caches(oiio):
    """b2r_trim_caches returns the band buffers a destroyed banded chain left in
    the per-device cache (and the plans, the tile slots); the library keeps working."""
    import torch
    lvl0 = synth.image(2048, 1024, 4, np.float32, seed=620)
    ch = oiio.MipChainBanded(oiio.ImageBuf(lvl0), [0, 0], filtername="box")
    want = ch.download(3)
    del ch
    torch.cuda.synchronize()
    free_cached, _ = torch.cuda.mem_get_info()
    oiio.trim_caches()
    free_trimmed, _ = torch.cuda.mem_get_info()
    assert free_trimmed >= free_cached + 2048 * 1024 * 16  # at least level 0's bands came back
    ch = oiio.MipChainBanded(oiio.ImageBuf(lvl0), [0, 0], filtername="box")
    assert np.array_equal(ch.download(3), want)
    out = np.zeros((512, 1024, 4), np.float32)
    assert oiio.ImageBufAlgo.resize(oiio.ImageBuf(out), oiio.ImageBuf(lvl0), filtername="lanczos3")
