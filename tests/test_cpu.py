"""CPU-side tests (no GPU): the oracle against the reference's golden vectors
and object code, the product's host logic (filters, error behaviour, fit
geometry), and that libb2r.so loads and exports everything include/b2r.h
declares.  No compute entry point is exercised here beyond its error paths."""
import ctypes as C
import os
import re
import subprocess

import numpy as np
import pytest

import synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


# ---------------------------------------------------------------- oracle ----
def test_oracle_filters_known_answer(oracle):
    """All 17 testsuite/filters/ref/<filter>.exr reproduced bit for bit."""
    G = np.load(os.path.join(GOLD, "filters_ref.npz"))
    L = oracle.lib()
    assert L.orc_filter_count() == 17
    for i in range(17):
        name = L.orc_filter_name(i).decode()
        delta = np.zeros((16, 16, 1), np.float32)
        delta[8, 8, 0] = 1
        dst = np.zeros((80, 80, 1), np.float32)
        oracle.resize(oracle.Img(dst), oracle.Img(delta), name)
        dst[0, i, 0] = 1.0
        assert np.array_equal(dst[..., 0].astype(np.float16), G[name]), name


def test_oracle_xform_goldens(oracle):
    X = np.load(os.path.join(GOLD, "xform_ref.npz"))
    gf = X["grid"].astype(np.float32) * np.float32(1 / 255.0)

    def rs(src, w, h):
        d = np.zeros((h, w, src.shape[2]), np.float32)
        oracle.resize(oracle.Img(d), oracle.Img(src))
        return oracle.quantize(d, np.uint8)

    for name, (w, h) in {"resize": (256, 256), "resize2": (250, 250),
                         "resize64": (64, 64)}.items():
        d = np.abs(rs(gf, w, h).astype(int) - X[name].astype(int))
        assert d.max() <= 1, name  # idiff -fail 0.004 == 1 LSB
    r64 = X["resize64"].astype(np.float32) * np.float32(1 / 255.0)
    d = np.abs(rs(r64, 512, 512).astype(int) - X["resize512"].astype(int))
    assert d.max() <= 1 and (d != 0).mean() < 1e-4
    # fit 360x240 of the square grid -> 240x240 letterboxed at x=60
    assert oracle.fit_geometry(1000, 1000, 360, 240) == (240, 240, 60, 0)
    assert oracle.fit_geometry(1000, 1000, 240, 360) == (240, 240, 0, 60)
    d = np.abs(rs(gf, 240, 240).astype(int) - X["fit"].astype(int))
    assert d.max() <= 3 and (d > 1).mean() < 2e-4  # idiff hardfail 0.012
    # resample goldens are exact
    for interp, name in [(True, "resample"), (False, "resample_nearest")]:
        d = np.zeros((128, 128, 4), np.float32)
        oracle.resample(oracle.Img(d), oracle.Img(gf), interp)
        assert np.array_equal(oracle.quantize(d, np.uint8), X[name]), name


def test_oracle_resized_offset_golden(oracle):
    """testsuite/oiiotool-xform resized-offset.exr: nonzero data origin."""
    X = np.load(os.path.join(GOLD, "xform_ref.npz"))
    tl, tr = np.array([1, 0, 0], np.float32), np.array([0, 1, 0], np.float32)
    bl, br = np.array([0, 0, 1], np.float32), np.array([0, 1, 1], np.float32)
    a = np.zeros((64, 64, 3), np.float32)
    for y in range(64):
        v = np.float32(y) / np.float32(63)
        for x in range(64):
            u = np.float32(x) / np.float32(63)
            a[y, x] = (tl * (1 - u) + tr * u) * (1 - v) + (bl * (1 - u) + br * u) * v
    dst = np.zeros((32, 32, 3), np.float32)
    oracle.resize(oracle.Img(dst, 50, 50, (0, 0, 128, 128)),
                  oracle.Img(a, 100, 100, (0, 0, 256, 256)))
    g = X["resized_offset"].astype(np.float32)
    assert np.abs(dst.astype(np.float16).astype(np.float32) - g).max() == 0.0


def test_oracle_rangecompress_goldens(oracle):
    """testsuite/oiiotool/ref/range{compress,expand}{,-luma}.tif: oiiotool reads
    the uint8 input as float, applies the op and writes uint8 (run.py:28-32)."""
    G = np.load(os.path.join(GOLD, "pixelmath_ref.npz"))

    def run(fn, u8, luma):
        f = u8.astype(np.float32) * np.float32(1 / 255.0)
        out = np.empty_like(f)
        fn(oracle.Img(out), oracle.Img(f), useluma=luma)
        return oracle.quantize(out, np.uint8)

    for luma, sfx in ((False, ""), (True, "_luma")):
        c = run(oracle.rangecompress, G["tahoe_small"], luma)
        assert np.array_equal(c, G["rangecompress" + sfx]), "rangecompress" + sfx
        e = run(oracle.rangeexpand, G["rangecompress" + sfx], luma)
        assert np.array_equal(e, G["rangeexpand" + sfx]), "rangeexpand" + sfx
    # in place, alpha left alone, luma refused when alpha is among the first three
    a = np.linspace(-2, 6, 4 * 5 * 7, dtype=np.float32).reshape(5, 7, 4)
    b = a.copy()
    oracle.rangecompress(oracle.Img(b), oracle.Img(b), useluma=True, alpha_channel=3)
    assert np.array_equal(b[..., 3], a[..., 3]) and not np.array_equal(b[..., :3], a[..., :3])
    c = a.copy()
    oracle.rangecompress(oracle.Img(c), oracle.Img(c), useluma=True, alpha_channel=1)
    d = a.copy()
    oracle.rangecompress(oracle.Img(d), oracle.Img(d), useluma=False, alpha_channel=1)
    assert np.array_equal(c, d)
    # expand undoes compress to float precision
    e = np.empty_like(a)
    oracle.rangeexpand(oracle.Img(e), oracle.Img(d), alpha_channel=1)
    assert np.allclose(e, a, rtol=2e-5, atol=1e-6)


def _ref_eval(R, name, w, h, mode, xs):
    out = np.empty_like(xs)
    sep = C.c_int()
    rc = R.ref_filter2d_eval(name.encode(), w, h, mode, len(xs), xs.ctypes.data,
                             xs.ctypes.data, out.ctypes.data, C.byref(sep),
                             None, None)
    assert rc == 0
    return out, sep.value


def test_oracle_and_product_filters_match_reference_object_code(oracle, oiio):
    """oracle/_ref/libref_filter.so is the reference's filter.cpp compiled
    unmodified; both restatements must agree with it to the bit."""
    R = oracle.ref_filter_lib()
    if R is None:
        pytest.skip("oracle/_ref not built (needs /root/reference)")
    L = oracle.lib()
    P = oiio.lib()
    xs = np.concatenate([np.linspace(-7, 7, 4001),
                         np.array([0, 1e-5, -1e-4, 0.5, 1, 2, 3, 1.5, -2.5])]
                        ).astype(np.float32)
    for i in range(R.ref_filter2d_count()):
        name = R.ref_filter2d_name(i).decode()
        for (w, h) in [(0.0, 0.0), (3.0, 3.0), (7.5, 3.25), (12.0, 24.0)]:
            f = P.b2r_filter2d_create(name.encode(), w, h)
            assert f
            for mode in (0, 1):
                ref, sep = _ref_eval(R, name, w, h, mode, xs)
                fo = L.orc_filter_xfilt if mode == 0 else L.orc_filter_yfilt
                fp = P.b2r_filter2d_xfilt if mode == 0 else P.b2r_filter2d_yfilt
                mine = np.array([fo(name.encode(), w, h, float(x)) for x in xs],
                                np.float32)
                prod = np.array([fp(f, float(x)) for x in xs], np.float32)
                assert np.array_equal(mine, ref), (name, w, h, mode)
                assert np.array_equal(prod, ref), (name, w, h, mode)
                assert sep == P.b2r_filter2d_separable(f)
            # 2-D evaluation on a diagonal
            ref2 = np.empty_like(xs)
            ys = (xs * np.float32(0.37)).astype(np.float32)
            R.ref_filter2d_eval(name.encode(), w, h, 2, len(xs), xs.ctypes.data,
                                ys.ctypes.data, ref2.ctypes.data, None, None, None)
            prod2 = np.array([P.b2r_filter2d_eval(f, float(x), float(y))
                              for x, y in zip(xs, ys)], np.float32)
            assert np.array_equal(prod2, ref2), (name, w, h)
            P.b2r_filter2d_destroy(f)


def test_filter_api_surface(oiio):
    P = oiio.lib()
    assert P.b2r_filter_count(2) == 17 and P.b2r_filter_count(1) == 15
    assert oiio.get_string_attribute("filter_list").split(";")[:3] == \
        ["box", "triangle", "gaussian"]
    # aliases are accepted by create() ...
    for alias, canon in [("catrom", "catmull-rom"), ("lanczos", "lanczos3"),
                         ("b-spline", "b-spline"), ("bspline", "b-spline"),
                         ("nuke-lanczos6", "lanczos3"),
                         ("sharp-gaussian", "gaussian")]:
        f = oiio.Filter2D.create(alias, 0.0)
        assert f is not None and f.name() == canon
    assert oiio.Filter2D.create("nope", 1.0) is None
    f = oiio.Filter2D.create("lanczos3", 0.0)
    assert f.width() == 6.0 and f.height() == 6.0 and f.separable()
    assert oiio.Filter2D.create("disk", 3.0).separable() is False
    # Filter1D: catmull-rom always reports width 4
    h = P.b2r_filter1d_create(b"catmull-rom", 8.0)
    assert P.b2r_filter1d_width(h) == 4.0
    P.b2r_filter1d_destroy(h)
    assert not P.b2r_filter1d_create(b"disk", 1.0)


# ------------------------------------------------------------- C ABI --------
def test_library_exports_every_declared_symbol(oiio):
    hdr = open(os.path.join(ROOT, "include", "b2r.h")).read()
    declared = set(re.findall(r"\b(b2r_[a-z0-9_]+)\s*\(", hdr))
    assert len(declared) >= 28
    out = subprocess.check_output(
        ["nm", "-D", "--defined-only",
         os.path.join(ROOT, "openimageio_b200", "libb2r.so")], text=True)
    exported = {l.split()[-1] for l in out.splitlines() if " T " in l}
    assert declared <= exported, declared - exported
    # nothing but the ABI leaks out
    assert all(s.startswith("b2r_") for s in exported), exported
    L = C.CDLL(os.path.join(ROOT, "openimageio_b200", "libb2r.so"))
    for s in declared:
        getattr(L, s)


def test_error_behaviour_without_touching_the_gpu(oiio):
    src = oiio.ImageBuf(synth.image(16, 16, 3, np.float32))
    # unknown filter name: the reference's message (xform.cpp:857); aliases
    # accepted by Filter2D::create are NOT accepted here
    for bad in ("catrom", "lanczos", "b-spline", "nope"):
        R = oiio.ImageBufAlgo.resize(src, bad, roi=oiio.ROI(0, 8, 0, 8, 0, 1, 0, 3))
        assert R.has_error
        assert R.geterror() == 'Filter "%s" not recognized' % bad
    R = oiio.ImageBufAlgo.resize(oiio.ImageBuf(), "box")
    assert R.has_error and R.geterror() == "Uninitialized input image"
    # channel mismatch the kernel cannot honour (dst wider than src)
    d = oiio.ImageBuf(np.zeros((8, 8, 4), np.float32))
    s = oiio.ImageBuf(np.zeros((16, 16, 3), np.float32))
    assert not oiio.ImageBufAlgo.resize(d, s)
    assert "channels" in d.geterror()
    # fit exact=1 is the warp path: refused, not silently approximated
    R = oiio.ImageBufAlgo.fit(src, exact=True, roi=oiio.ROI(0, 8, 0, 8, 0, 1, 0, 3))
    assert R.has_error


def test_no_gpu_means_error_not_fallback(oiio):
    if oiio.lib().b2r_device_count() > 0:
        pytest.skip("a GPU is present")
    src = oiio.ImageBuf(synth.image(16, 16, 3, np.float32))
    R = oiio.ImageBufAlgo.resize(src, "box", roi=oiio.ROI(0, 8, 0, 8, 0, 1, 0, 3))
    assert R.has_error and "no CPU fallback" in R.geterror()
    with pytest.raises(oiio.B2RError):
        oiio.MipChain(np.zeros((8, 8, 4), np.float32))


def test_fit_geometry_matches_oracle(oiio, oracle):
    P = oiio.lib()
    rng = np.random.RandomState(7)
    for _ in range(300):
        sw, sh, fw, fh = [int(v) for v in rng.randint(1, 4000, 4)]
        for mode in ("letterbox", "width", "height", "bogus"):
            v = [C.c_int() for _ in range(4)]
            P.b2r_fit_geometry(sw, sh, fw, fh, mode.encode(),
                               *[C.byref(i) for i in v])
            assert tuple(i.value for i in v) == \
                oracle.fit_geometry(sw, sh, fw, fh, mode)


def test_product_does_not_touch_the_oracle():
    """The product path must not import, link or execute oracle/."""
    pkg = os.path.join(ROOT, "openimageio_b200")
    for dirpath, _, files in os.walk(pkg):
        if os.path.basename(dirpath) == "build":
            continue
        for f in files:
            if f.endswith((".py", ".cpp", ".cu", ".h", "Makefile")):
                txt = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in txt.lower() or f == "__init__.py" and \
                    "oracle" not in txt, (dirpath, f)
    out = subprocess.check_output(["ldd", os.path.join(pkg, "libb2r.so")], text=True)
    assert "oracle" not in out


def test_mip_level_sizes(oracle):
    """write_mipmap's level rule: floor halving until 1x1."""
    lv = oracle.mip_chain(np.zeros((5, 12, 1), np.float32), "box")
    assert [(a.shape[1], a.shape[0]) for a in lv] == [(6, 2), (3, 1), (1, 1)]


# ------------------------------------------------ oracle: warp / rotate ----
def _fit_pattern(w, h):
    """oiiotool --pattern constant:color=.25,.25,.25,1 WxH 4 --box:color=1,1,1
    0,0,W-1,H-1 --box:color=1,0,0 4,4,W-5,H-5 -d half (oiiotool-xform/run.py:73-78)."""
    a = np.zeros((h, w, 4), np.float32)
    a[..., :3] = 0.25
    a[..., 3] = 1

    def box(x1, y1, x2, y2, col):
        a[y1, x1:x2 + 1] = col
        a[y2, x1:x2 + 1] = col
        a[y1:y2 + 1, x1] = col
        a[y1:y2 + 1, x2] = col
    box(0, 0, w - 1, h - 1, (1, 1, 1, 1))
    box(4, 4, w - 5, h - 5, (1, 0, 0, 1))
    return a.astype(np.float16)


def test_oracle_warp_goldens(oracle):
    """testsuite/oiiotool-xform: --warp, --rotate, --resize:from/to and
    --fit:exact=1 known answers, all reproduced BIT FOR BIT.  (rotated.tif and
    warped.tif were made from the resize.tif of the same run, which on the
    machine that made them equals the oracle's resize of grid.tif, not the
    stored ref/resize.tif -- so the source here is the oracle's own resize.)"""
    import math
    X = np.load(os.path.join(GOLD, "xform_ref.npz"))
    W = np.load(os.path.join(GOLD, "warp_ref.npz"))
    gf = X["grid"].astype(np.float32) * np.float32(1 / 255.0)
    d = np.zeros((256, 256, 4), np.float32)
    oracle.resize(oracle.Img(d), oracle.Img(gf))
    rs = oracle.quantize(d, np.uint8)
    src = oracle.Img(rs)
    M = [0.7071068, 0.7071068, 0, -0.7071068, 0.7071068, 0, 128, -53.01933, 1]
    out = np.zeros_like(rs)
    oracle.warp(oracle.Img(out), src, M)
    assert np.array_equal(out, W["warped"])
    ang = np.float32(45.0) * np.float32(math.pi / 180.0)  # oiiotool.cpp --rotate
    for name, (cx, cy) in {"rotated": (128.0, 128.0),
                           "rotated_offcenter": (50.0, 50.0)}.items():
        out = np.zeros_like(rs)
        oracle.warp(oracle.Img(out), src, oracle.rotate_matrix(ang, cx, cy),
                    wrap="black")
        assert np.array_equal(out, W[name]), name
    G = oracle.Img(X["grid"].copy())
    for name, to in [("resizefrom", (0, 0, 64, 64)), ("resizefromto", (0, 0, 32, 32)),
                     ("resizefromtooffset", (5, -5, 32, 32))]:
        Mf = oracle.fromto_matrix((300, 300, 200, 200), to)
        out = np.zeros((64, 64, 4), np.uint8)
        oracle.warp(oracle.Img(out, 0, 0, (0, 0, 64, 64)), G, Mf)
        assert np.array_equal(out, W[name]), name
    # --fit:exact=1
    Mx, rw, rh = oracle.fit_exact_matrix(288, 216, 216, 162)
    out = np.zeros((162, 216, 3), np.float16)
    oracle.warp(oracle.Img(out), oracle.Img(W["target1"].copy()), Mx,
                "blackman-harris", 3.0, 3.0, wrap="black", edgeclamp=True)
    assert np.array_equal(out.view(np.uint16), W["fit4"].view(np.uint16))
    for name, (sw, sh), size, mode in [("fitw_letterbox_200", (256, 128), 200, "letterbox"),
                                       ("fith_width_300", (128, 256), 300, "width")]:
        Mx, rw, rh = oracle.fit_exact_matrix(sw, sh, size, size, mode)
        wr, hr = np.float32(rw) / np.float32(sw), np.float32(rh) / np.float32(sh)
        fn, bw = ("blackman-harris", 3.0) if (wr > 1 or hr > 1) else ("lanczos3", 6.0)
        fw = np.float32(bw) * max(np.float32(1), wr)
        fh = np.float32(bw) * max(np.float32(1), hr)
        out = np.zeros((size, size, 4), np.float16)
        oracle.warp(oracle.Img(out), oracle.Img(_fit_pattern(sw, sh)), Mx, fn, fw, fh,
                    wrap="black", edgeclamp=True)
        assert np.array_equal(out.view(np.uint16), W[name].view(np.uint16)), name


def test_oracle_warp_matrix_algebra(oracle):
    M = oracle.rotate_matrix(0.3, 10.0, 20.0)
    Mi = oracle.m33_inverse(M)
    assert np.allclose(M.astype(np.float64) @ Mi.astype(np.float64), np.eye(3), atol=1e-5)
    # transform(M, roi): bounding box of the pixel-centre corners
    T = np.array([[1, 0, 0], [0, 1, 0], [10, -5, 1]], np.float32)
    assert oracle.transform_roi(T, (0, 100, 0, 50)) == (10, 110, -5, 45)


# ------------------------------------- oracle: make_kernel / convolve ------
def test_oracle_convolve_goldens(oracle):
    """testsuite/oiiotool: --kernel bspline 15x15 --convolve (bit for bit),
    --blur 5x5 and --unsharp (1 LSB on < 2e-5 of the values: the stored images
    come from a build whose float sums round differently in a handful of
    places; idiff's tolerance is 1 LSB everywhere)."""
    P = np.load(os.path.join(GOLD, "pixelmath_ref.npz"))
    Cv = np.load(os.path.join(GOLD, "conv_ref.npz"))
    t = P["tahoe_small"]
    K, known = oracle.make_kernel("bspline", 15, 15)
    assert known and K.shape == (15, 15) and abs(K.sum() - 1) < 1e-6
    out = np.zeros_like(t)
    oracle.convolve(oracle.Img(out), oracle.Img(t), K)
    assert np.array_equal(out, Cv["bspline_blur"])
    K, _ = oracle.make_kernel("gaussian", 5, 5)
    out = np.zeros_like(t)
    oracle.convolve(oracle.Img(out), oracle.Img(t), K)
    d = np.abs(out.astype(int) - Cv["gauss5x5_blur"].astype(int))
    assert d.max() <= 1 and (d != 0).mean() < 2e-5
    out = np.zeros_like(t)
    oracle.unsharp_mask(oracle.Img(out), oracle.Img(t))
    d = np.abs(out.astype(int) - Cv["unsharp"].astype(int))
    assert d.max() <= 1 and (d != 0).mean() < 2e-5


def test_oracle_make_kernel_shapes(oracle):
    K, known = oracle.make_kernel("gaussian", 3.0, 3.0)
    assert known and K.shape == (3, 3) and K[1, 1] == K.max()
    K, known = oracle.make_kernel("gaussian", 1, 5.0)       # unsharp width > 3
    assert K.shape == (5, 1)
    K, known = oracle.make_kernel("binomial", 5, 5)
    assert known and np.allclose(K[2] * 256, np.array([6, 24, 36, 24, 6]))
    K, known = oracle.make_kernel("laplacian", 3, 3)
    assert known and K.sum() == 0 and K[1, 1] == -4
    K, known = oracle.make_kernel("nonesuch", 4, 4)
    assert not known and K.shape == (5, 5) and np.allclose(K, 1 / 25)


def test_oracle_st_warp_golden(oracle):
    """testsuite/oiiotool-xform st_warped.tif: <= 1 LSB on < 2e-4 of the values
    (the ST field goes through powf, whose last bit is platform dependent)."""
    import ctypes as C
    X = np.load(os.path.join(GOLD, "xform_ref.npz"))
    W = np.load(os.path.join(GOLD, "warp_ref.npz"))
    gf = X["grid"].astype(np.float32) * np.float32(1 / 255.0)
    d = np.zeros((256, 256, 4), np.float32)
    oracle.resize(oracle.Img(d), oracle.Img(gf))
    rs = oracle.quantize(d, np.uint8)
    libm = C.CDLL("libm.so.6")
    libm.powf.restype = C.c_float
    libm.powf.argtypes = [C.c_float, C.c_float]
    u = np.arange(256, dtype=np.float32) / np.float32(255)
    g = np.array([libm.powf(float(v), 1.2) for v in u], np.float32)
    st = np.zeros((256, 256, 3), np.float32)
    st[..., 0] = g[None, :]
    st[..., 1] = g[:, None]
    out = np.zeros_like(rs)
    oracle.st_warp(oracle.Img(out), oracle.Img(rs), oracle.Img(st))
    dd = np.abs(out.astype(int) - W["st_warped"].astype(int))
    assert dd.max() <= 1 and (dd != 0).mean() < 2e-4


def test_oracle_fix_nonfinite(oracle):
    """fixNonFinite_ (imagebufalgo_pixelmath.cpp:1424-1500): counts, black,
    box3 with clipped neighbourhoods, half sums in half precision."""
    a = np.arange(5 * 6 * 2, dtype=np.float32).reshape(5, 6, 2)
    a[2, 3, 0] = np.nan
    a[0, 0, 1] = np.inf
    a[4, 5, :] = -np.inf
    d = np.zeros_like(a)
    assert oracle.fix_nonfinite(oracle.Img(d), oracle.Img(a), "box3") == 3
    nb = [a[j, i, 0] for j in (1, 2, 3) for i in (2, 3, 4) if (j, i) != (2, 3)]
    assert d[2, 3, 0] == np.float32(sum(nb)) / np.float32(8)
    assert d[0, 0, 1] == (a[0, 1, 1] + a[1, 0, 1] + a[1, 1, 1]) / np.float32(3)
    d = np.zeros_like(a)
    assert oracle.fix_nonfinite(oracle.Img(d), oracle.Img(a), "black") == 3
    assert d[2, 3, 0] == 0 and d[4, 5, 0] == 0 and d[4, 5, 1] == 0 and d[1, 1, 1] == a[1, 1, 1]
    d = np.zeros_like(a)
    assert oracle.fix_nonfinite(oracle.Img(d), oracle.Img(a), "none") == 3
    assert np.isnan(d[2, 3, 0])
    h = (a * 100).astype(np.float16)
    d = np.zeros_like(h)
    oracle.fix_nonfinite(oracle.Img(d), oracle.Img(h), "box3")
    s = np.float16(0)
    for j in (1, 2, 3):
        for i in (2, 3, 4):
            if (j, i) != (2, 3):
                s = np.float16(np.float32(s) + np.float32(h[j, i, 0]))
    assert d[2, 3, 0] == np.float16(np.float32(s) / np.float32(8))


def test_oracle_fillholes_golden(oracle):
    """testsuite/oiiotool src/hole.tif --fillholes == ref/tahoe-filled.tif, bit
    for bit (triangle resize pyramid + divide_by_alpha + over)."""
    G = np.load(os.path.join(GOLD, "fillholes_ref.npz"))
    out = oracle.fillholes_pushpull(G["hole"], 3)
    assert np.array_equal(out, G["tahoe_filled"])


def _oiio_hash(a, extra="", blocksize=0):
    """computePixelHashSHA1 restated with hashlib (the oracle for the product's
    own SHA-1): imagebufalgo_compare.cpp:817-892."""
    import hashlib
    h = a.shape[0]
    if blocksize <= 0 or blocksize >= h:
        s = hashlib.sha1(np.ascontiguousarray(a).tobytes())
        s.update(extra.encode())
        return s.hexdigest().upper()
    s = hashlib.sha1()
    for y in range(0, h, blocksize):
        s.update(hashlib.sha1(np.ascontiguousarray(a[y:y + blocksize]).tobytes())
                 .hexdigest().upper().encode())
    s.update(extra.encode())
    return s.hexdigest().upper()


def test_pixel_hash_sha1_reference_goldens(oiio):
    """The two SHA-1 values the reference's own test suite holds for
    testsuite/common/grid.tif (testsuite/maketx/ref/out.txt:4, 20): the plain
    pixel hash of the uint8 image (`SHA-1:` of oiiotool --info) and maketx's
    oiio:SHA-1 -- 256-row blocks of the float32 top level with the filter name
    "box " appended (maketexture.cpp:1909-1934).  Host images: no GPU needed."""
    X = np.load(os.path.join(GOLD, "xform_ref.npz"))
    grid = X["grid"]
    assert grid.shape == (1000, 1000, 4) and grid.dtype == np.uint8
    IBA = oiio.ImageBufAlgo
    assert IBA.computePixelHashSHA1(grid) == "97D6548FF9E9BFD21424B773B1D84FACDF0C1797"
    top = grid.astype(np.float32) * np.float32(1.0 / 255.0)   # convert_type<uint8,float>
    assert (IBA.computePixelHashSHA1(top, "box ", blocksize=256)
            == "9B24D4CC05313A43973AC384718D81D19183B691")
    # and the hashlib restatement agrees with both
    assert _oiio_hash(grid) == "97D6548FF9E9BFD21424B773B1D84FACDF0C1797"
    assert _oiio_hash(top, "box ", 256) == "9B24D4CC05313A43973AC384718D81D19183B691"


@pytest.mark.parametrize("shape,dt,blocksize", [((37, 53, 3), np.float32, 0),
                                                ((700, 129, 4), np.uint16, 256),
                                                ((256, 64, 1), np.uint8, 256),
                                                ((513, 64, 2), np.float16, 256),
                                                ((1, 1, 1), np.uint8, 0),
                                                ((1000, 31, 3), np.float32, 7)])
def test_pixel_hash_sha1_matches_hashlib(oiio, shape, dt, blocksize):
    a = synth.image(shape[1], shape[0], shape[2], dt, seed=4242)
    for extra in ("", "lanczos3 sharpen_A=0.5 highlightcomp=1 "):
        for nt in (1, 5):
            got = oiio.ImageBufAlgo.computePixelHashSHA1(a, extra, blocksize=blocksize,
                                                         nthreads=nt)
            assert got == _oiio_hash(a, extra, blocksize)
    # a region of interest hashes only its columns / rows
    if shape[0] > 20 and shape[1] > 20:
        roi = oiio.ROI(3, shape[1] - 5, 2, shape[0] - 7, 0, 1, 0, shape[2])
        got = oiio.ImageBufAlgo.computePixelHashSHA1(a, "x", roi=roi, blocksize=blocksize)
        assert got == _oiio_hash(a[2:shape[0] - 7, 3:shape[1] - 5], "x", blocksize)


# ---- the TIFF checker of the tile-output tests, pinned against libtiff ------
@pytest.mark.parametrize("big", [False, True])
@pytest.mark.parametrize("predictor", [0, 3])
@pytest.mark.parametrize("nch", [1, 3, 4])
def test_tiff_mini_against_libtiff(tmp_path, nch, predictor, big):
    """tests/tiff_mini.py (the numpy TIFF reader / writer the GPU tests use to
    check b2r_tileset_write_tiff) against libtiff itself, through OpenCV / PIL:
    a file it writes -- tiled, deflate, Predictor 3 as libtiff's fpDiff defines
    it, classic and BigTIFF, one directory per level -- decodes to the same
    pixels in both readers."""
    import tiff_mini
    rng = np.random.default_rng(40 + nch)
    levels = [rng.standard_normal((h, w, nch)).astype(np.float32)
              for h, w in ((150, 200), (75, 100), (1, 1))]
    path = tmp_path / "m.tif"
    tiff_mini.write(path, levels, tile=64, predictor=predictor, big=big)
    mine = tiff_mini.read(path)
    assert len(mine) == 3
    for a, b in zip(levels, mine):
        assert np.array_equal(a, b["pixels"]) and b["padding_is_zero"]
        assert b["big"] == big
    theirs = tiff_mini.read_libtiff(path, nch)
    if theirs is None:
        pytest.skip("neither OpenCV nor PIL here")
    assert len(theirs) == 3
    for a, b in zip(levels, theirs):
        assert np.array_equal(a, b)
    # half samples: only the numpy reader decodes them
    halves = [a.astype(np.float16) for a in levels]
    tiff_mini.write(path, halves, tile=32, predictor=predictor, big=big)
    for a, b in zip(halves, tiff_mini.read(path)):
        assert np.array_equal(a, b["pixels"])


@pytest.mark.parametrize("dt", [np.uint8, np.uint16])
@pytest.mark.parametrize("predictor", [0, 2])
def test_tiff_mini_integers_against_libtiff(tmp_path, dt, predictor):
    """The integer side of tests/tiff_mini.py (uint8 / uint16 samples, TIFF's
    horizontal predictor as libtiff's horDiff8 / horDiff16 define it) against
    libtiff through OpenCV."""
    import tiff_mini
    rng = np.random.default_rng(50)
    for nch in (1, 3, 4):
        levels = [rng.integers(0, np.iinfo(dt).max + 1, (h, w, nch)).astype(dt)
                  for h, w in ((150, 200), (75, 100), (1, 1))]
        path = tmp_path / "i.tif"
        tiff_mini.write(path, levels, tile=64, predictor=predictor, big=(nch == 3))
        for a, b in zip(levels, tiff_mini.read(path)):
            assert np.array_equal(a, b["pixels"]) and b["padding_is_zero"]
        theirs = tiff_mini.read_libtiff(path, nch)
        if theirs is None:
            pytest.skip("no OpenCV here")
        assert len(theirs) == 3 and all(np.array_equal(a, b) for a, b in zip(levels, theirs))


@pytest.mark.parametrize("dt", [np.float32, np.float16])
def test_exr_mini_against_openexr(tmp_path, dt):
    """tests/exr_mini.py (the numpy OpenEXR reader / writer the GPU tests use to
    check every MIP level of b2r_tileset_write_exr's files) against the OpenEXR
    library itself, through OpenCV: a tiled, MIP-mapped, zip-compressed file it
    writes decodes to the same level 0 there, and its own reader returns every
    level."""
    import exr_mini
    rng = np.random.default_rng(60)
    for nch, (w, h), tile in ((1, (200, 150), 64), (3, (64, 64), 64), (4, (130, 97), 32)):
        levels = [rng.standard_normal((hh, ww, nch)).astype(dt)
                  for ww, hh in exr_mini.level_sizes(w, h)]
        path = tmp_path / "m.exr"
        exr_mini.write(path, levels, tile=tile)
        got, attrs = exr_mini.read(path)
        assert len(got) == len(levels)
        assert all(np.array_equal(a, b) for a, b in zip(levels, got))
        l0 = exr_mini.read_openexr_level0(path, nch)
        if l0 is None:
            pytest.skip("no OpenCV / OpenEXR here")
        assert np.array_equal(l0, levels[0].astype(np.float32))


def struct_version(path):
    import struct
    with open(path, "rb") as f:
        return struct.unpack("<II", f.read(8))[1]


def test_exr_and_tiff_writers_on_host(tmp_path):
    """b2r_tileset_write_exr / b2r_tileset_write_tiff are host code: fed a hand-built tile set of stored
    tiles (tests/cpp/exr_harness.cpp) it must produce a file whose every level
    reads back through exr_mini and whose level 0 the OpenEXR library (OpenCV)
    decodes bit for bit -- float and half, 1-4 channels, ragged and degenerate
    shapes.  Needs the objects of the library build (build() leaves them)."""
    import glob
    import exr_mini
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    csrc = os.path.join(root, "openimageio_b200", "csrc")
    objs = [o for o in glob.glob(os.path.join(csrc, "build", "*.o"))
            if not o.endswith("b2r_api_tiles.o")]
    nvcc = "/usr/local/cuda/bin/nvcc"
    if len(objs) < 10 or not os.path.exists(nvcc):
        pytest.skip("library objects or nvcc not here")
    exe, obj = str(tmp_path / "exr_harness"), str(tmp_path / "exr_harness.o")
    subprocess.check_call(["g++", "-std=c++17", "-O1", "-I", os.path.join(root, "include"),
                           "-I", csrc, "-I", "/usr/local/cuda/include", "-c",
                           os.path.join(root, "tests", "cpp", "exr_harness.cpp"), "-o", obj])
    subprocess.check_call([nvcc, "-o", exe, obj] + objs
                          + ["-lcudart_static", "-lpthread", "-ldl", "-lrt", "-lz"],
                          stderr=subprocess.DEVNULL)

    def val(lv, h, w, c):
        y, x, cc = np.meshgrid(np.arange(h), np.arange(w), np.arange(c), indexing="ij")
        return (((x * 7 + y * 13 + cc * 17 + lv * 29) % 1000).astype(np.float32)
                / np.float32(1000.0) - np.float32(0.25))

    path = str(tmp_path / "a.exr")
    for W, H, nch, half, tile in ((200, 150, 4, 0, 64), (200, 150, 4, 1, 64), (130, 97, 3, 0, 32),
                                  (64, 64, 1, 1, 64), (41, 70, 2, 1, 16), (1, 1, 3, 0, 16),
                                  (513, 2, 4, 0, 64)):
        subprocess.check_call([exe, path, str(W), str(H), str(nch), str(half), str(tile)],
                              stdout=subprocess.DEVNULL)
        levels, attrs = exr_mini.read(path)
        sizes = exr_mini.level_sizes(W, H)
        assert len(levels) == len(sizes)
        for lv, (w, h) in enumerate(sizes):
            want = val(lv, h, w, nch)
            want = want.astype(np.float16) if half else want
            assert levels[lv].dtype == want.dtype and np.array_equal(levels[lv], want)
        assert attrs["oiio:SHA-1"] == ("string", b"0123ABCD")
        if nch != 2:
            l0 = exr_mini.read_openexr_level0(path, nch)
            if l0 is not None:
                w0 = val(0, H, W, nch)
                w0 = w0.astype(np.float16).astype(np.float32) if half else w0
                assert np.array_equal(l0, w0)
    # a long attribute name switches the file to the long-names version flag; the
    # OpenEXR library still takes it
    subprocess.check_call([exe, path, "100", "60", "3", "0", "32", "long"],
                          stdout=subprocess.DEVNULL)
    levels, attrs = exr_mini.read(path)
    assert attrs["an_attribute_name_of_more_than_31_characters"] == ("string", b"x")
    assert struct_version(path) == (2 | 0x200 | 0x400)
    l0 = exr_mini.read_openexr_level0(path, 3)
    if l0 is not None:
        assert np.array_equal(l0, val(0, 60, 100, 3))
    # the TIFF container around the same hand-built (stored) tiles: libtiff reads
    # the float files, the numpy reader all of them; classic and BigTIFF
    import tiff_mini
    tpath = str(tmp_path / "a.tif")
    for W, H, nch, half, tile, big in ((200, 150, 4, 0, 64, 0), (130, 97, 3, 0, 32, 1),
                                       (200, 150, 4, 1, 64, 0), (41, 70, 1, 0, 16, 1)):
        subprocess.check_call([exe, tpath, str(W), str(H), str(nch), str(half), str(tile),
                               str(big)], stdout=subprocess.DEVNULL)
        dirs = tiff_mini.read(tpath)
        sizes = exr_mini.level_sizes(W, H)
        assert len(dirs) == len(sizes)
        for lv, (w, h) in enumerate(sizes):
            want = val(lv, h, w, nch)
            want = want.astype(np.float16) if half else want
            assert np.array_equal(dirs[lv]["pixels"], want) and dirs[lv]["padding_is_zero"]
            assert dirs[lv]["big"] == bool(big) and dirs[lv]["tags"][259] == [1]
            assert dirs[lv]["tags"][270] == "oiio:SHA-1=0123ABCD"
        if not half:
            theirs = tiff_mini.read_libtiff(tpath, nch)
            if theirs is not None:
                assert len(theirs) == len(sizes)
                assert all(np.array_equal(a, d["pixels"]) for a, d in zip(theirs, dirs))


def test_python_mirror_struct_layouts_match_the_header(oiio, tmp_path):
    """The ctypes structures of the Python mirror against include/b2r.h as the C
    compiler lays it out: same size for every structure that crosses the ABI
    (a field added on one side only would shift everything behind it)."""
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    pairs = [("b2r_image", oiio._Image), ("b2r_roi", oiio._Roi), ("b2r_options", oiio._Options),
             ("b2r_report", oiio.Report), ("b2r_pixel_stats_t", oiio._PixelStatsC),
             ("b2r_mip_options", oiio._MipOptions), ("b2r_tile_options", oiio._TileOptions),
             ("b2r_tiff_options", oiio._TiffOptions), ("b2r_exr_options", oiio._ExrOptions)]
    src = tmp_path / "sizes.c"
    src.write_text('#include <stdio.h>\n#include "b2r.h"\nint main(void){\n'
                   + "".join('printf("%s %%zu\\n", sizeof(%s));\n' % (n, n) for n, _ in pairs)
                   + "return 0;}\n")
    exe = str(tmp_path / "sizes")
    subprocess.check_call(["gcc", "-I", os.path.join(root, "include"), str(src), "-o", exe])
    sizes = dict(l.split() for l in subprocess.check_output([exe], text=True).splitlines())
    for name, cls in pairs:
        assert int(sizes[name]) == C.sizeof(cls), (name, sizes[name], C.sizeof(cls))
