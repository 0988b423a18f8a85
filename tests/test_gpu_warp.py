"""GPU parity for the warp path (ImageBufAlgo::warp / rotate / fit exact=1 /
oiiotool --resize:from/to): the CUDA kernel through the C ABI against the CPU
oracle on identical inputs, and against the reference's own golden images.

Bars: uint8/uint16 +-1 LSB (and almost everywhere exact), float 1e-5 relative,
half 1 ulp.  The footprint arithmetic is bit-identical by construction; only
sinf/cosf inside the filter profiles can differ from glibc's."""
import math
import os

import numpy as np
import pytest

import synth
from test_gpu_parity import check, _np_dtype

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _buf(oiio, arr, win=None):
    B = oiio.ImageBuf(arr)
    if win:
        x, y, full = win
        B.spec().x, B.spec().y = x, y
        (B.spec().full_x, B.spec().full_y, B.spec().full_width,
         B.spec().full_height) = full if full else (x, y, arr.shape[1], arr.shape[0])
    return B


def warp_both(oiio, oracle, src, dst_shape, dst_dtype, M, filtername="", fw=0.0,
              wrap="default", edgeclamp=False, src_win=None, dst_win=None,
              roi=None, prefill=None, device_resident=False):
    h, w = dst_shape
    c = src.shape[2]
    g = np.zeros((h, w, c), dst_dtype) if prefill is None else prefill.copy()
    o = g.copy()
    sx, sy, sfull = src_win if src_win else (0, 0, None)
    dx, dy, dfull = dst_win if dst_win else (0, 0, None)
    # oracle: filter resolved the way get_warp_filter does
    name = filtername or "lanczos3"
    oracle.warp(oracle.Img(o, dx, dy, dfull), oracle.Img(src, sx, sy, sfull), M,
                name, fw, fw, wrap, edgeclamp, roi)
    S = _buf(oiio, src, src_win)
    D = _buf(oiio, g, dst_win)
    r = oiio.ROI(roi[0], roi[1], roi[2], roi[3], 0, 1, 0, c) if roi else None
    if device_resident:
        import torch

        def up(a):
            if a.dtype == np.uint16:
                return torch.from_numpy(a.view(np.int16)).cuda().view(torch.uint16)
            return torch.from_numpy(a).cuda()
        S.pixels, D.pixels = up(src), up(g)
    ok = oiio.ImageBufAlgo.warp(D, S, M, filtername=filtername, filterwidth=fw,
                                wrap=wrap, edgeclamp=edgeclamp, roi=r)
    assert ok, D.geterror()
    if device_resident:
        g = D.get_pixels()
    return g, o


def _rot(oiio, deg, cx, cy):
    return oiio.rotate_matrix(np.float32(deg) * np.float32(math.pi / 180.0), cx, cy)


PERSPECTIVE = np.array([[0.9, 0.12, 0.0006], [-0.2, 1.05, -0.0004],
                        [9.0, -7.0, 1.0]], np.float32)


def test_warp_goldens(oiio, oracle):
    """oiiotool-xform warped.tif / rotated*.tif / resizefrom*.tif through the
    CUDA path."""
    X = np.load(os.path.join(GOLD, "xform_ref.npz"))
    W = np.load(os.path.join(GOLD, "warp_ref.npz"))
    gf = X["grid"].astype(np.float32) * np.float32(1 / 255.0)
    d = np.zeros((256, 256, 4), np.float32)
    oracle.resize(oracle.Img(d), oracle.Img(gf))
    rs = oracle.quantize(d, np.uint8)  # see test_cpu.test_oracle_warp_goldens
    IBA = oiio.ImageBufAlgo
    M = [0.7071068, 0.7071068, 0, -0.7071068, 0.7071068, 0, 128, -53.01933, 1]
    out = IBA.warp(oiio.ImageBuf(rs), M)
    assert not out.has_error, out.geterror()
    dd = np.abs(out.pixels.astype(int) - W["warped"].astype(int))
    assert dd.max() <= 1 and (dd != 0).mean() < 1e-3
    ang = np.float32(45.0) * np.float32(math.pi / 180.0)
    out = IBA.rotate(oiio.ImageBuf(rs), ang)
    dd = np.abs(out.pixels.astype(int) - W["rotated"].astype(int))
    assert dd.max() <= 1 and (dd != 0).mean() < 1e-3
    out = IBA.rotate(oiio.ImageBuf(rs), ang, 50.0, 50.0)
    dd = np.abs(out.pixels.astype(int) - W["rotated_offcenter"].astype(int))
    assert dd.max() <= 1 and (dd != 0).mean() < 1e-3
    G = oiio.ImageBuf(X["grid"].copy())
    for name, to in [("resizefrom", None), ("resizefromto", (0, 0, 32, 32)),
                     ("resizefromtooffset", (5, -5, 32, 32))]:
        out = oiio.oiiotool_resize(G, 64, 64, from_geom=(300, 300, 200, 200),
                                   to_geom=to)
        assert not out.has_error, out.geterror()
        dd = np.abs(out.pixels.astype(int) - W[name].astype(int))
        assert dd.max() <= 1 and (dd != 0).mean() < 1e-3, name


def test_fit_exact_goldens(oiio, oracle):
    from test_cpu import _fit_pattern
    W = np.load(os.path.join(GOLD, "warp_ref.npz"))
    IBA = oiio.ImageBufAlgo
    out = IBA.fit(oiio.ImageBuf(W["target1"].copy()), filtername="blackman-harris",
                  exact=True, roi=oiio.ROI(0, 216, 0, 162, 0, 1, 0, 3))
    assert not out.has_error, out.geterror()
    check(out.pixels, W["fit4"])
    for name, (sw, sh), size, mode in [("fitw_letterbox_200", (256, 128), 200, "letterbox"),
                                       ("fith_width_300", (128, 256), 300, "width")]:
        out = IBA.fit(oiio.ImageBuf(_fit_pattern(sw, sh)), fillmode=mode, exact=True,
                      roi=oiio.ROI(0, size, 0, size, 0, 1, 0, 4))
        assert not out.has_error, out.geterror()
        assert out.pixels.shape == (size, size, 4)
        check(out.pixels, W[name])


@pytest.mark.parametrize("dt,st", [("float", "float"), ("uint8", "uint8"),
                                   ("uint16", "uint16"), ("half", "half"),
                                   ("float", "uint8"), ("uint8", "float"),
                                   ("half", "uint16"), ("uint16", "half")])
@pytest.mark.parametrize("nch", [1, 3, 4])
def test_warp_type_pairs(oiio, oracle, dt, st, nch):
    src = synth.image(97, 61, nch, _np_dtype(st), seed=11 + nch)
    g, o = warp_both(oiio, oracle, src, (70, 90), _np_dtype(dt),
                     _rot(oiio, 17.0, 40.0, 30.0))
    check(g, o)
    if g.dtype in (np.uint8, np.uint16):
        assert (g != o).mean() < 2e-3


@pytest.mark.parametrize("filt", ["box", "triangle", "gaussian", "sharp-gaussian",
                                  "catmull-rom", "blackman-harris", "sinc", "lanczos3",
                                  "radial-lanczos3", "nuke-lanczos6", "mitchell",
                                  "bspline", "disk", "cubic", "keys", "simon", "rifman"])
def test_warp_every_filter(oiio, oracle, filt):
    src = synth.image(64, 48, 3, np.float32, seed=5)
    g, o = warp_both(oiio, oracle, src, (48, 64), np.float32, PERSPECTIVE, filt)
    check(g, o)


@pytest.mark.parametrize("wrap", ["default", "black", "clamp", "periodic", "mirror"])
@pytest.mark.parametrize("cropped", [False, True])
def test_warp_wrap_modes(oiio, oracle, wrap, cropped):
    """Footprints that leave the image on every side, with the data window
    equal to the display window and as a crop inside / overscan outside it."""
    src = synth.image(40, 30, 4, np.float32, seed=9)
    win = (7, -3, (0, 0, 50, 32)) if cropped else None
    M = np.array([[0.5, 0.05, 0], [-0.04, 0.45, 0], [20, 14, 1]], np.float32)
    g, o = warp_both(oiio, oracle, src, (64, 96), np.float32, M, "mitchell",
                     wrap=wrap, src_win=win, dst_win=(-8, -6, (0, 0, 80, 52)))
    check(g, o)


@pytest.mark.parametrize("st", ["float", "uint8"])
def test_warp_edgeclamp_and_minify(oiio, oracle, st):
    """edgeclamp (fit exact / --resize:edgeclamp) and a 3.7x minification whose
    footprint (23 columns) is wider than the kernel's weight cache... and one
    wider still."""
    src = synth.image(200, 160, 3, _np_dtype(st), seed=3)
    for scale in (0.27, 0.11):
        M = np.array([[scale, 0, 0], [0, scale, 0], [3.3, 2.1, 1]], np.float32)
        g, o = warp_both(oiio, oracle, src, (50, 64), _np_dtype(st), M, "lanczos3",
                         wrap="black", edgeclamp=True)
        check(g, o)


def test_warp_roi_channels_and_prefill(oiio, oracle):
    """A dst ROI smaller than dst and a 5-channel image (two channel passes):
    pixels outside the ROI keep their values."""
    src = synth.image(50, 40, 5, np.float32, seed=21)
    pre = synth.image(60, 44, 5, np.float32, seed=22)
    g, o = warp_both(oiio, oracle, src, (44, 60), np.float32,
                     _rot(oiio, -30.0, 25.0, 20.0), "catmull-rom",
                     roi=(5, 41, 3, 40), prefill=pre)
    check(g, o)
    assert np.array_equal(g[:3], pre[:3]) and np.array_equal(g[:, :5], pre[:, :5])


def test_warp_device_resident(oiio, oracle):
    for st in ("float", "half", "uint16", "uint8"):
        src = synth.image(128, 96, 4, _np_dtype(st), seed=31)
        g, o = warp_both(oiio, oracle, src, (96, 128), _np_dtype(st), PERSPECTIVE,
                         device_resident=True)
        check(g, o)


def test_warp_recompute_roi_and_errors(oiio, oracle):
    IBA = oiio.ImageBufAlgo
    src = oiio.ImageBuf(synth.image(32, 24, 3, np.float32, seed=1))
    M = _rot(oiio, 45.0, 16.0, 12.0)
    out = IBA.warp(src, M, recompute_roi=True)
    assert not out.has_error
    r = oracle.transform_roi(M, (0, 32, 0, 24))
    assert (out.spec().x, out.spec().x + out.spec().width, out.spec().y,
            out.spec().y + out.spec().height) == r
    ref = np.zeros(out.pixels.shape, np.float32)
    oracle.warp(oracle.Img(ref, r[0], r[2], (0, 0, 32, 24)), oracle.Img(src.pixels), M)
    check(out.pixels, ref)
    bad = IBA.warp(src, M, filtername="catrom")
    assert bad.has_error and bad.geterror() == 'Filter "catrom" not recognized'
    assert IBA.warp(oiio.ImageBuf(), oiio.ImageBuf(), M) is False


def test_warp_singular_matrix(oiio, oracle):
    """A singular matrix inverts to identity (Imath's non-throwing inverse)."""
    src = synth.image(20, 20, 3, np.float32, seed=2)
    M = np.zeros((3, 3), np.float32)
    g, o = warp_both(oiio, oracle, src, (20, 20), np.float32, M, "triangle")
    check(g, o)


# ---------------------------------------------------------------- st_warp ----
def _st_identity_gamma(n=256, gamma=1.2):
    """--pattern fill:topleft=0,0,0:topright=1,0,0:bottomleft=0,1,0:
    bottomright=1,1,0 NxN 3 --powc 1.2 (oiiotool-xform/run.py:99-101), with
    glibc's powf like the reference."""
    import ctypes as C
    libm = C.CDLL("libm.so.6")
    libm.powf.restype = C.c_float
    libm.powf.argtypes = [C.c_float, C.c_float]
    u = np.arange(n, dtype=np.float32) / np.float32(n - 1)
    one = np.float32(1)
    ramp = (one - u) * np.float32(0) + u          # bilerp along one axis
    g = np.array([libm.powf(float(v), gamma) for v in ramp], np.float32)
    st = np.zeros((n, n, 3), np.float32)
    st[..., 0] = g[None, :]
    st[..., 1] = g[:, None]
    return st


def test_st_warp_golden(oiio, oracle):
    X = np.load(os.path.join(GOLD, "xform_ref.npz"))
    W = np.load(os.path.join(GOLD, "warp_ref.npz"))
    gf = X["grid"].astype(np.float32) * np.float32(1 / 255.0)
    d = np.zeros((256, 256, 4), np.float32)
    oracle.resize(oracle.Img(d), oracle.Img(gf))
    rs = oracle.quantize(d, np.uint8)
    st = _st_identity_gamma()
    ref = np.zeros_like(rs)
    oracle.st_warp(oracle.Img(ref), oracle.Img(rs), oracle.Img(st))
    dd = np.abs(ref.astype(int) - W["st_warped"].astype(int))
    assert dd.max() <= 1 and (dd != 0).mean() < 2e-4           # oracle pinned
    out = oiio.ImageBufAlgo.st_warp(oiio.ImageBuf(rs), oiio.ImageBuf(st))
    assert not out.has_error, out.geterror()
    dd = np.abs(out.pixels.astype(int) - W["st_warped"].astype(int))
    assert dd.max() <= 1 and (dd != 0).mean() < 1e-3
    dd = np.abs(out.pixels.astype(int) - ref.astype(int))
    assert dd.max() <= 1 and (dd != 0).mean() < 1e-3


@pytest.mark.parametrize("st_type", ["float", "half", "uint16"])
@pytest.mark.parametrize("filt", ["lanczos3", "triangle", "disk"])
def test_st_warp_parity(oiio, oracle, st_type, filt):
    """Smooth ST field, a source with 3 channels, flips, non-default ST
    channels, a dst twice the source's size (filter radius scales with
    1 / scale).  For the positive filters the field leaves the image on every
    side (black taps, clipped footprints).  lanczos3 stays inside: where a
    clipped footprint keeps only a few taps of a filter with negative lobes the
    weight sum can nearly cancel, and sum / total amplifies the last-bit
    difference between the device's and glibc's sinf a thousandfold -- on the
    CPU just as well."""
    src = synth.image(48, 40, 3, np.float32, seed=41)
    h, w = 80, 96
    yy, xx = np.mgrid[0:h, 0:w].astype(np.float32)
    a, b = (0.76, 0.12) if filt == "lanczos3" else (1.2, -0.1)
    s = (xx / w * a + b + 0.03 * np.sin(yy / 7)).astype(np.float32)
    t = (yy / h * a + b + 0.03 * np.cos(xx / 5)).astype(np.float32)
    if st_type != "float":
        s, t = np.clip(s, 0, 1), np.clip(t, 0, 1)
    st32 = np.stack([np.zeros_like(s), t, s], -1)            # chan_s=2, chan_t=1
    st = (oracle.quantize(st32, np.uint16) if st_type == "uint16"
          else st32.astype(_np_dtype(st_type)))
    ref = np.zeros((h, w, 3), np.float32)
    oracle.st_warp(oracle.Img(ref), oracle.Img(src), oracle.Img(st), filt, chan_s=2,
                   chan_t=1, flip_t=True)
    got = np.zeros_like(ref)
    D = oiio.ImageBuf(got)
    ok = oiio.ImageBufAlgo.st_warp(D, oiio.ImageBuf(src), oiio.ImageBuf(st),
                                   filtername=filt, chan_s=2, chan_t=1, flip_t=True)
    assert ok, D.geterror()
    check(got, ref)


def test_st_warp_errors(oiio):
    IBA = oiio.ImageBufAlgo
    src = oiio.ImageBuf(np.zeros((8, 8, 3), np.float32))
    st = oiio.ImageBuf(np.zeros((8, 8, 2), np.float32))
    assert IBA.st_warp(src, st, chan_t=2).geterror() == \
        "ImageBufAlgo::st_warp : Out-of-range T channel index: 2"
    assert IBA.st_warp(src, oiio.ImageBuf()).geterror() == \
        "ImageBufAlgo::st_warp : Uninitialized ST buffer"
    assert IBA.st_warp(src, st, filtername="nope").geterror() == \
        'Filter "nope" not recognized'
    far = oiio.ImageBuf(np.zeros((8, 8, 2), np.float32))
    far.spec().x = 100
    assert IBA.st_warp(src, far).geterror() == \
        "ImageBufAlgo::st_warp : Output ROI does not intersect ST buffer."


@pytest.mark.gpu
@pytest.mark.parametrize("st", ["float", "uint8", "half", "uint16"])
def test_warp_tolerance_mode(oiio, oracle, st):
    """b2r_options::contract_fma: the accumulation uses packed fused multiply-adds;
    the footprint arithmetic is unchanged, so the result stays within north_star's
    tolerance of the oracle (1e-5 of the value range for float, half an ulp more
    for half, 1 LSB for integers) -- and it is not the default."""
    dt = {"float": np.float32, "uint8": np.uint8, "half": np.float16, "uint16": np.uint16}[st]
    src = synth.image(240, 160, 4, dt, seed=501)
    M = oracle.rotate_matrix(0.4, 120.0, 80.0)
    ref = np.zeros((160, 240, 4), dt)
    oracle.warp(oracle.Img(ref), oracle.Img(src), M, "lanczos3", wrap="black")
    exact = np.zeros_like(ref)
    assert oiio.ImageBufAlgo.warp(oiio.ImageBuf(exact), oiio.ImageBuf(src), M,
                                  filtername="lanczos3", wrap="black")
    fast = np.zeros_like(ref)
    oiio.ImageBufAlgo.contract_fma = True
    try:
        assert oiio.ImageBufAlgo.warp(oiio.ImageBuf(fast), oiio.ImageBuf(src), M,
                                      filtername="lanczos3", wrap="black")
    finally:
        oiio.ImageBufAlgo.contract_fma = False
    if st in ("uint8", "uint16"):
        assert np.abs(fast.astype(int) - ref.astype(int)).max() <= 1
        assert (fast != ref).mean() < 1e-2
    else:
        tol = 1e-5 if st == "float" else 2.0 ** -10
        assert np.abs(fast.astype(np.float64) - ref.astype(np.float64)).max() <= tol
    # a different rounding sequence really ran (float: some last bits differ)
    if st == "float":
        assert not np.array_equal(fast, exact)
