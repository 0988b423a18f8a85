"""GPU parity: the CUDA path (through the C ABI, via the host mirror of
ImageBufAlgo) against the CPU oracle on identical seeded inputs.

Bars (BASELINE.json north_star): uint8/uint16 bit-exact or +-1 LSB;
half/float max abs error <= 1e-5 relative (to the value range of the image);
the exact kernel is held to bit-exactness for float outputs.
"""
import os

import numpy as np
import pytest

import synth

pytestmark = pytest.mark.gpu

REL_TOL = 1e-5


def _np_dtype(name):
    return {"uint8": np.uint8, "uint16": np.uint16, "half": np.float16,
            "float": np.float32}[name]


def run_both(oiio, oracle, src_arr, dst_shape, dst_dtype, filtername="",
             filterwidth=0.0, src_win=None, dst_win=None, roi=None,
             path=0, prefill=None):
    """Returns (gpu_out, oracle_out, report).  *_win = (x, y, full tuple)."""
    h, w = dst_shape
    c = src_arr.shape[2]
    g = np.zeros((h, w, c), dst_dtype) if prefill is None else prefill.copy()
    o = g.copy()
    sx, sy, sfull = src_win if src_win else (0, 0, None)
    dx, dy, dfull = dst_win if dst_win else (0, 0, None)
    # oracle
    osrc = oracle.Img(src_arr, sx, sy, sfull)
    odst = oracle.Img(o, dx, dy, dfull)
    oracle.resize(odst, osrc, filtername, filterwidth, roi)
    # product
    S = oiio.ImageBuf(src_arr)
    S.spec().x, S.spec().y = sx, sy
    if sfull:
        (S.spec().full_x, S.spec().full_y, S.spec().full_width,
         S.spec().full_height) = sfull
    else:
        S.spec().full_x, S.spec().full_y = sx, sy
    D = oiio.ImageBuf(g)
    D.spec().x, D.spec().y = dx, dy
    if dfull:
        (D.spec().full_x, D.spec().full_y, D.spec().full_width,
         D.spec().full_height) = dfull
    else:
        D.spec().full_x, D.spec().full_y = dx, dy
    r = None
    if roi:
        r = oiio.ROI(roi[0], roi[1], roi[2], roi[3], 0, 1, 0, c)
    ok = oiio.ImageBufAlgo.resize(D, S, filtername=filtername,
                                  filterwidth=filterwidth, roi=r, path=path)
    assert ok, D.geterror()
    return g, o, oiio.ImageBufAlgo.last_report


def check(g, o, exact=False, lsb_tol=1):
    if g.dtype in (np.uint8, np.uint16):
        d = np.abs(g.astype(np.int64) - o.astype(np.int64))
        assert d.max() <= (0 if exact else lsb_tol), "max LSB diff %d" % d.max()
        return d.max()
    gf, of = g.astype(np.float64), o.astype(np.float64)
    assert np.array_equal(np.isnan(gf), np.isnan(of))
    m = ~np.isnan(of)
    assert np.array_equal(np.isinf(gf[m]), np.isinf(of[m]))
    inf = np.isinf(of)
    assert np.array_equal(gf[inf], of[inf])  # +Inf / -Inf where the reference has them
    fin = np.isfinite(of)
    if exact and g.dtype == np.float32:
        assert np.array_equal(g[fin], o[fin])
        return 0.0
    scale = max(1.0, np.abs(of[fin]).max()) if fin.any() else 1.0
    err = np.abs(gf[fin] - of[fin]).max() / scale if fin.any() else 0.0
    tol = REL_TOL if g.dtype == np.float32 else 2.0 ** -10  # half: 1 ulp
    assert err <= tol, "rel err %g" % err
    return err


# ---- the ten native (dst, src) type pairs --------------------------------
PAIRS = [("float", "float"), ("float", "uint8"), ("float", "half"),
         ("float", "uint16"), ("uint8", "float"), ("uint8", "uint8"),
         ("half", "float"), ("half", "half"), ("uint16", "float"),
         ("uint16", "uint16")]


@pytest.mark.parametrize("dt,st", PAIRS)
@pytest.mark.parametrize("nch", [1, 3, 4])
def test_type_pairs_downsize(oiio, oracle, dt, st, nch):
    src = synth.image(201, 157, nch, _np_dtype(st), seed=11)
    g, o, rep = run_both(oiio, oracle, src, (75, 96), _np_dtype(dt))
    check(g, o)
    assert rep.path_used == oiio.PATH_FUSED


@pytest.mark.parametrize("dt,st", [("float", "float"), ("uint8", "uint8"),
                                   ("half", "half"), ("uint16", "uint16")])
def test_exact_kernel_bitexact(oiio, oracle, dt, st):
    src = synth.image(97, 83, 3, _np_dtype(st), seed=5)
    g, o, rep = run_both(oiio, oracle, src, (40, 51), _np_dtype(dt),
                         path=oiio.PATH_EXACT)
    check(g, o, exact=True)
    assert rep.path_used == oiio.PATH_EXACT


FILTERS = ["box", "triangle", "gaussian", "sharp-gaussian", "catmull-rom",
           "blackman-harris", "sinc", "lanczos3", "radial-lanczos3",
           "nuke-lanczos6", "mitchell", "bspline", "disk", "cubic", "keys",
           "simon", "rifman"]


@pytest.mark.parametrize("filt", FILTERS)
def test_all_filters_down_and_up(oiio, oracle, filt):
    src = synth.image(64, 48, 4, np.float32, seed=3)
    for shape in [(24, 32), (100, 130)]:
        g, o, rep = run_both(oiio, oracle, src, shape, np.float32, filt)
        if filt in ("radial-lanczos3", "disk"):
            assert rep.path_used == oiio.PATH_EXACT
            check(g, o)  # device sinf differs from glibc by ulps
        else:
            check(g, o)


def test_filters_known_answer_golden(oiio):
    """testsuite/filters/run.py: delta 16x16 -> 80x80 per filter, half."""
    import os
    G = np.load(os.path.join(os.path.dirname(__file__), "golden",
                             "filters_ref.npz"))
    names = oiio.get_string_attribute("filter_list").split(";")
    assert len(names) == 17
    for i, f in enumerate(names):
        delta = np.zeros((16, 16, 1), np.float32)
        delta[8, 8, 0] = 1.0
        S = oiio.ImageBuf(delta)
        R = oiio.ImageBufAlgo.resize(S, f, roi=oiio.ROI(0, 80, 0, 80, 0, 1, 0, 1))
        assert not R.has_error, R.geterror()
        out = R.get_pixels()[..., 0].copy()
        out[0, i] = 1.0
        h = out.astype(np.float16).astype(np.float32)
        ref = G[f].astype(np.float32)
        assert np.abs(h - ref).max() <= 2.0 ** -11, f  # <= 1 half ulp near 1


@pytest.mark.parametrize("name,shape,tol", [("resize", (256, 256), 1),
                                            ("resize2", (250, 250), 1),
                                            ("resize64", (64, 64), 1)])
def test_grid_goldens(oiio, name, shape, tol):
    """testsuite/oiiotool-xform: grid.tif --resize; goldens are uint8 written
    from a float resize, matched to 1 LSB (idiff 0.004)."""
    import os
    X = np.load(os.path.join(os.path.dirname(__file__), "golden", "xform_ref.npz"))
    gf = X["grid"].astype(np.float32) * np.float32(1 / 255.0)
    S = oiio.ImageBuf(gf)
    R = oiio.ImageBufAlgo.resize(S, roi=oiio.ROI(0, shape[1], 0, shape[0], 0, 1, 0, 4))
    assert not R.has_error, R.geterror()
    import oracle_py
    q = oracle_py.quantize(R.get_pixels(), np.uint8)
    d = np.abs(q.astype(int) - X[name].astype(int))
    assert d.max() <= tol


def test_windows_offset_and_crop(oiio, oracle):
    """Nonzero origins, data window inside the full window (black outside),
    dst data window smaller than its full window, sub-ROI."""
    src = synth.image(64, 64, 3, np.float32, seed=9)
    g, o, _ = run_both(oiio, oracle, src, (32, 32), np.float16,
                       src_win=(100, 100, (0, 0, 256, 256)),
                       dst_win=(50, 50, (0, 0, 128, 128)))
    check(g, o)
    # dst window reaching outside the source data -> black taps + clamp
    g, o, _ = run_both(oiio, oracle, src, (60, 60), np.float32,
                       src_win=(100, 100, (0, 0, 256, 256)),
                       dst_win=(40, 40, (0, 0, 128, 128)))
    check(g, o)
    # sub-ROI leaves the rest of dst untouched
    pre = synth.image(90, 70, 3, np.float32, seed=1)
    src2 = synth.image(180, 140, 3, np.float32, seed=2)
    g, o, _ = run_both(oiio, oracle, src2, (70, 90), np.float32,
                       roi=(10, 55, 5, 61), prefill=pre)
    check(g, o)
    assert np.array_equal(g[:5], pre[:5]) and np.array_equal(g[:, :10], pre[:, :10])


def test_overscan_source_uses_exact(oiio, oracle):
    """Data window larger than the display window: WrapClamp is not
    separable, the exact kernel must take it."""
    src = synth.image(80, 72, 4, np.float32, seed=4)
    g, o, rep = run_both(oiio, oracle, src, (30, 34), np.float32,
                         src_win=(-8, -4, (0, 0, 64, 64)),
                         dst_win=(0, 0, (0, 0, 34, 30)))
    check(g, o, exact=True)
    assert rep.path_used == oiio.PATH_EXACT


@pytest.mark.parametrize("shape", [(1, 1), (1, 37), (53, 1), (7, 5), (300, 2)])
def test_ragged_sizes(oiio, oracle, shape):
    src = synth.image(37, 29, 4, np.float32, seed=8)
    g, o, _ = run_both(oiio, oracle, src, shape, np.float32)
    check(g, o)
    src = synth.image(3, 2, 3, np.uint8, seed=8)
    g, o, _ = run_both(oiio, oracle, src, shape, np.uint8)
    check(g, o)


def test_extreme_ratio_wide_footprint(oiio, oracle):
    """grid 1000 -> 64: 95 taps per axis (testsuite resize64)."""
    src = synth.image(500, 400, 4, np.uint8, seed=6)
    g, o, _ = run_both(oiio, oracle, src, (26, 32), np.uint8)
    check(g, o)


def test_upsample_mitchell_catrom_uint8(oiio, oracle):
    """BASELINE config 3 at reduced size: RGB uint8 x4, +-1 LSB."""
    src = synth.image(120, 68, 3, np.uint8, seed=12)
    for f in ("mitchell", "catmull-rom"):
        g, o, rep = run_both(oiio, oracle, src, (272, 480), np.uint8, f)
        check(g, o)
        assert rep.path_used == oiio.PATH_FUSED


EXPAND_SHAPES = [  # (src h, w) -> (dst h, w)
    ((68, 120), (272, 480)),    # 4x both axes
    ((51, 77), (102, 154)),     # 2x, ragged source width
    ((40, 63), (60, 95)),       # ~1.5x, widths not multiples of 4
    ((33, 50), (99, 150)),      # 3x
    ((30, 200), (120, 100)),    # up 4x in y, down 2x in x
    ((45, 31), (45, 124)),      # same height, 4x in x
    ((9, 6), (300, 17)),        # tiny source, tall destination
]


@pytest.mark.parametrize("nch", [1, 2, 3, 4])
@pytest.mark.parametrize("dt,st", [("uint8", "uint8"), ("float", "float"),
                                   ("half", "half"), ("uint16", "uint16"),
                                   ("uint8", "float"), ("float", "uint16")])
def test_expand_kernel_upsizing(oiio, oracle, nch, dt, st):
    """The upsizing variant of the fused kernel (vmode < 0): every channel
    count, source/destination type, layout choice and window width."""
    used = set()
    for i, (sshape, dshape) in enumerate(EXPAND_SHAPES):
        src = synth.image(sshape[1], sshape[0], nch, _np_dtype(st), seed=40 + i)
        if st == "float":
            # HDR and negative values: the quantiser's clamps must engage
            src = (src * np.float32(3.0) - np.float32(0.7)).astype(np.float32)
        for f in ("", "mitchell", "lanczos3", "triangle"):
            if i >= 2 and f not in ("", "mitchell"):
                continue
            g, o, rep = run_both(oiio, oracle, src, dshape, _np_dtype(dt), f)
            check(g, o)
            assert rep.path_used == oiio.PATH_FUSED
            used.add(rep.fused_vmode < 0)
    assert True in used  # the expand kernel did run


def test_expand_kernel_nonfinite_and_windows(oiio, oracle):
    src = synth.image(52, 40, 4, np.float32, seed=77)
    src[10, 17, 1] = np.inf
    src[30, 40, 2] = np.nan
    g, o, rep = run_both(oiio, oracle, src, (160, 208), np.float32, "mitchell")
    # the exact kernel re-does the cells that hold a non-finite output (the
    # NaN / Inf pattern must equal the reference's); the rest stays fused
    check(g, o)
    assert rep.nonfinite_redo == 1
    # ROI inside a larger destination, offset windows, unaligned ROI origin
    src = synth.image(52, 40, 3, np.uint8, seed=78)
    pre = synth.image(170, 130, 3, np.uint8, seed=79)
    g, o, rep = run_both(oiio, oracle, src, (130, 170), np.uint8, "catmull-rom",
                         src_win=(3, 5, (3, 5, 52, 40)),
                         dst_win=(7, 2, (7, 2, 170, 130)),
                         roi=(18, 150, 11, 120), prefill=pre)
    check(g, o)
    assert rep.fused_vmode < 0


def test_nonfinite_zero_weight_skip(oiio, oracle):
    """Inf/NaN under zero-weight taps must not poison the output
    (imagebufalgo_xform.cpp:749-759 skips zero weights).  The stream kernel
    resolves it in place (predicated accumulate in the V pass, re-filtered
    windows in the H pass); the fused kernel marks the cells that hold a
    non-finite output and the exact kernel re-does those cells."""
    src = synth.image(96, 80, 4, np.float32, seed=13)
    src[10, 17, 1] = np.inf
    src[40, 60, 2] = np.nan
    src[70, 5, 0] = -np.inf
    g, o, rep = run_both(oiio, oracle, src, (40, 48), np.float32)
    check(g, o)  # includes the NaN / +Inf / -Inf pattern
    assert rep.path_used == oiio.PATH_FUSED
    if rep.nonfinite_redo == 1:
        # the re-done cells are the exact kernel's: bit-exact around the defects
        assert np.array_equal(g[0:16, 0:48], o[0:16, 0:48], equal_nan=True)
    # ratio 1 with lanczos3: integer taps are exact zeros around the centre
    g, o, rep = run_both(oiio, oracle, src, (80, 96), np.float32, "triangle")
    check(g, o)


def test_explicit_filterwidth_and_errors(oiio, oracle):
    src = synth.image(64, 64, 3, np.float32, seed=2)
    g, o, _ = run_both(oiio, oracle, src, (40, 40), np.float32, "gaussian", 5.0)
    check(g, o)
    S = oiio.ImageBuf(src)
    R = oiio.ImageBufAlgo.resize(S, "catrom", roi=oiio.ROI(0, 8, 0, 8, 0, 1, 0, 3))
    assert R.has_error and R.geterror() == 'Filter "catrom" not recognized'
    R = oiio.ImageBufAlgo.resize(oiio.ImageBuf(), "box")
    assert R.has_error and "Uninitialized input image" in R.geterror()


def test_device_resident_tensors(oiio, oracle):
    """src and dst as torch CUDA tensors: no host copies, stream ordered."""
    import torch
    src = synth.image(256, 192, 4, np.float32, seed=21)
    o = np.zeros((96, 128, 4), np.float32)
    oracle.resize(oracle.Img(o), oracle.Img(src))
    ts = torch.from_numpy(src).cuda()
    td = torch.zeros((96, 128, 4), dtype=torch.float32, device="cuda")
    ok = oiio.ImageBufAlgo.resize(oiio.ImageBuf(td), oiio.ImageBuf(ts))
    assert ok
    torch.cuda.synchronize()
    check(td.cpu().numpy(), o)


def test_resample(oiio, oracle):
    for st, dt in [(np.uint8, np.uint8), (np.float32, np.float32),
                   (np.float16, np.float32)]:
        src = synth.image(100, 80, 4, st, seed=31)
        for interp in (True, False):
            for shape in [(37, 45), (160, 210)]:
                o = np.zeros(shape + (4,), dt)
                oracle.resample(oracle.Img(o), oracle.Img(src), interp)
                g = np.zeros(shape + (4,), dt)
                ok = oiio.ImageBufAlgo.resample(oiio.ImageBuf(g), oiio.ImageBuf(src),
                                                interpolate=interp)
                assert ok
                check(g, o, exact=True)
    # cropped source -> scalar path with WrapClamp / black
    src = synth.image(40, 40, 3, np.float32, seed=32)
    o = np.zeros((64, 64, 3), np.float32)
    oracle.resample(oracle.Img(o, full=(0, 0, 64, 64)),
                    oracle.Img(src, 30, 30, (0, 0, 128, 128)), True)
    S = oiio.ImageBuf(src)
    S.spec().x = S.spec().y = 30
    S.spec().full_x = S.spec().full_y = 0
    S.spec().full_width = S.spec().full_height = 128
    g = np.zeros((64, 64, 3), np.float32)
    assert oiio.ImageBufAlgo.resample(oiio.ImageBuf(g), S, interpolate=True)
    check(g, o, exact=True)


def test_resample_goldens(oiio):
    import os
    import oracle_py
    X = np.load(os.path.join(os.path.dirname(__file__), "golden", "xform_ref.npz"))
    gf = X["grid"].astype(np.float32) * np.float32(1 / 255.0)
    for interp, name in [(True, "resample"), (False, "resample_nearest")]:
        R = oiio.ImageBufAlgo.resample(oiio.ImageBuf(gf), interp,
                                       roi=oiio.ROI(0, 128, 0, 128, 0, 1, 0, 4))
        assert not R.has_error, R.geterror()
        q = oracle_py.quantize(R.get_pixels(), np.uint8)
        assert np.array_equal(q, X[name])


def test_fit(oiio, oracle):
    src = synth.image(200, 100, 3, np.float32, seed=41)
    for (fw, fh), mode in [((90, 60), "letterbox"), ((60, 90), "letterbox"),
                           ((90, 60), "width"), ((90, 60), "height")]:
        R = oiio.ImageBufAlgo.fit(oiio.ImageBuf(src), fillmode=mode,
                                  roi=oiio.ROI(0, fw, 0, fh, 0, 1, 0, 3))
        assert not R.has_error, R.geterror()
        rw, rh, xo, yo = oracle.fit_geometry(200, 100, fw, fh, mode)
        sp = R.spec()
        assert (sp.width, sp.height, sp.x, sp.y) == (rw, rh, xo, yo)
        assert (sp.full_width, sp.full_height) == (fw, fh)
        o = np.zeros((rh, rw, 3), np.float32)
        oracle.resize(oracle.Img(o), oracle.Img(src))
        check(R.get_pixels(), o)


@pytest.mark.parametrize("filt", ["box", "lanczos3", "triangle"])
@pytest.mark.parametrize("size", [(64, 64), (100, 37), (33, 130)])
def test_mip_chain(oiio, oracle, filt, size):
    """write_mipmap level loop, device resident; every level against the
    oracle's level loop (levels >= 1 are not pinned by the reference)."""
    w, h = size
    lvl0 = synth.image(w, h, 4, np.float32, seed=51)
    ref = oracle.mip_chain(lvl0, filt)
    chain = oiio.MipChain(lvl0, filt)
    assert chain.nlevels == len(ref) + 1
    for i, r in enumerate(ref):
        assert chain.level_size(i + 1) == (r.shape[1], r.shape[0])
        g = chain.download(i + 1)
        if filt == "box":
            assert np.array_equal(g, r), "level %d" % (i + 1)
        else:
            # errors compound level to level; both sides take their own
            # previous level as input
            assert np.abs(g - r).max() <= 1e-5 * (i + 1), "level %d" % (i + 1)


def test_banded_multi_device_equals_single(oiio, oracle):
    """Row bands on several devices (here the same GPU listed three times:
    the host-side split, band+halo staging and disjoint stores are what is
    under test) reproduce the single-call result bit for bit."""
    src = synth.image(300, 260, 4, np.float32, seed=61)
    one = np.zeros((130, 150, 4), np.float32)
    assert oiio.ImageBufAlgo.resize(oiio.ImageBuf(one), oiio.ImageBuf(src))
    many = np.zeros_like(one)
    D = oiio.ImageBuf(many)
    assert oiio.resize_banded(D, oiio.ImageBuf(src), [0, 0, 0]), D.geterror()
    assert np.array_equal(one, many)
    ref = np.zeros_like(one)
    oracle.resize(oracle.Img(ref), oracle.Img(src))
    check(many, ref)


FULL = {
    # name: (src w, h, nch, dtype, dst w, h, filter)  -- BASELINE.json configs
    "c0": (7680, 4320, 4, np.float32, 3840, 2160, "lanczos3"),
    "c2": (7680, 4320, 4, np.float16, 3840, 2160, "blackman-harris"),
    "c3": (1920, 1080, 3, np.uint8, 7680, 4320, "mitchell"),
    "c5": (3840, 2160, 3, np.uint16, 1920, 1080, "lanczos3"),
}


FULL["c1"] = (1920, 1080, 3, np.float32, 960, 540, "lanczos3")
FULL["c3cr"] = (1920, 1080, 3, np.uint8, 7680, 4320, "catmull-rom")


def _check_blocks(out, ref, block=270):
    """check() in row blocks (keeps the float64 temporaries small); returns the
    worst error."""
    worst = 0
    for y in range(0, out.shape[0], block):
        worst = max(worst, check(out[y:y + block], ref[y:y + block]))
    return worst


@pytest.mark.parametrize("name", sorted(FULL))
def test_full_size_configs_whole_image_against_oracle(oiio, oracle, name):
    """BASELINE.json's full sizes (C0, C1, C2, C3 with mitchell AND catmull-rom,
    C5): EVERY pixel of the GPU image is compared with the oracle (threaded
    over the host cores; SURVEY 8d: +-1 LSB for integers over the whole image,
    1e-5 of the value range for float, one half ulp for half) -- strip and band
    seams of the launch plan included, nothing sampled.  The host-memory call
    (band-pipelined) and the device-resident call (one launch) are both checked."""
    import torch
    sw, sh, c, dt, dw, dh, filt = FULL[name]
    src = synth.image(sw, sh, c, dt, seed=20261019)
    ref = np.zeros((dh, dw, c), dt)
    oracle.resize(oracle.Img(ref), oracle.Img(src), filt, nthreads=os.cpu_count() or 1)
    # host buffers through the public API
    out = np.zeros((dh, dw, c), dt)
    assert oiio.ImageBufAlgo.resize(oiio.ImageBuf(out), oiio.ImageBuf(src), filtername=filt)
    assert oiio.ImageBufAlgo.last_report.path_used == oiio.PATH_FUSED
    _check_blocks(out, ref)
    # device-resident: the whole frame in one launch
    tdt = {np.float32: torch.float32, np.uint8: torch.uint8, np.float16: torch.float16,
           np.uint16: torch.uint16}[dt]
    st = torch.from_numpy(src.view(np.int16) if dt == np.uint16 else src)
    st = (st.view(torch.uint16) if dt == np.uint16 else st).cuda()
    dd = torch.zeros((dh, dw, c), dtype=tdt, device="cuda")
    assert oiio.ImageBufAlgo.resize(oiio.ImageBuf(dd), oiio.ImageBuf(st), filtername=filt)
    got = dd.cpu()
    got = (got.view(torch.int16).numpy().view(np.uint16) if dt == np.uint16 else got.numpy())
    _check_blocks(got, ref)


@pytest.mark.parametrize("filt,size", [("box", 4096), ("lanczos3", 4096), ("box", (4100, 3000)),
                                       ("lanczos3", (5000, 4097))])
def test_mip_chain_large_per_level_against_oracle(oiio, oracle, filt, size):
    """The C4 chain at >= 4096^2 (the 16K^2 texture itself would
    keep the O(taps^2) oracle busy for minutes): EVERY level is compared with
    the oracle's level loop over the whole level -- box bit-exact, lanczos3
    within 1e-5 of the value range of the level -- and odd sizes take the
    bilinear resize_block path."""
    w, h = (size, size) if isinstance(size, int) else size
    src = synth.image(w, h, 4, np.float32, seed=20261023)
    levels = oracle.mip_chain(src, filt)  # [level 1, level 2, ...]
    ch = oiio.MipChain(src, filtername=filt)
    assert ch.nlevels == len(levels) + 1
    for lv in range(1, ch.nlevels):
        got, ref = ch.download(lv), levels[lv - 1]
        assert got.shape == ref.shape
        if filt == "box":
            assert np.array_equal(got, ref), "box level %d" % lv
        else:
            # each level is computed from the previous GPU level in the product and
            # from the previous oracle level in the oracle: errors of <= 1e-5 per
            # level add up at most linearly (values are in [0, 1))
            err = np.abs(got.astype(np.float64) - ref).max()
            assert err <= 1e-5 * lv, "level %d: %g" % (lv, err)


@pytest.mark.parametrize("dt,nch", [(np.float32, 4), (np.uint8, 3), (np.float16, 4),
                                    (np.uint16, 3), (np.float32, 1)])
def test_device_strided_rows_and_canaries(oiio, oracle, dt, nch):
    """Device-resident src/dst whose rows are padded (ystride > width) and
    which sit inside larger allocations: results match the oracle and not one
    byte outside the dst ROI rows is written (compute-sanitizer is closed on
    this pool, so out-of-bounds stores are hunted with canaries)."""
    import torch
    tdt = {np.float32: torch.float32, np.uint8: torch.uint8,
           np.float16: torch.float16, np.uint16: torch.uint16}[dt]
    sw, sh, dw, dh = 211, 150, 97, 71
    src = synth.image(sw, sh, nch, dt, seed=123)
    ref = np.zeros((dh, dw, nch), dt)
    oracle.resize(oracle.Img(ref), oracle.Img(src))
    spad, dpad = 5, 7  # extra pixels per row
    S = torch.zeros((sh + 4, sw + spad, nch), dtype=tdt, device="cuda")
    host = torch.from_numpy(src.view(np.int16) if dt == np.uint16 else src)
    if dt == np.uint16:
        host = host.view(torch.uint16)
    S[2:2 + sh, :sw] = host.cuda()
    fill = 77 if dt in (np.uint8, np.uint16) else 0.5
    D = torch.full((dh + 6, dw + dpad, nch), fill, dtype=tdt, device="cuda")
    sview, dview = S[2:2 + sh, :sw], D[3:3 + dh, :dw]
    ok = oiio.ImageBufAlgo.resize(oiio.ImageBuf(dview), oiio.ImageBuf(sview))
    assert ok
    torch.cuda.synchronize()
    Dh = D.cpu()
    if dt == np.uint16:
        Dn = Dh.view(torch.int16).numpy().view(np.uint16)
    else:
        Dn = Dh.numpy()
    check(Dn[3:3 + dh, :dw], ref)
    canary = np.full_like(Dn, fill)
    canary[3:3 + dh, :dw] = Dn[3:3 + dh, :dw]
    assert np.array_equal(canary, Dn), "bytes outside the dst ROI were written"


def test_concurrent_calls_from_threads(oiio, oracle):
    """The ABI is re-entrant: concurrent calls on distinct dst images (same
    shape -> shared cached plan, and different shapes) give the same results
    as serial calls."""
    import threading
    srcs = [synth.image(160, 120, 4, np.float32, seed=200 + i) for i in range(6)]
    shapes = [(60, 80), (60, 80), (45, 70), (60, 80), (90, 33), (45, 70)]
    outs = [np.zeros(s + (4,), np.float32) for s in shapes]
    errs = []

    def work(i):
        for _ in range(5):
            D = oiio.ImageBuf(outs[i])
            if not oiio.ImageBufAlgo.resize(D, oiio.ImageBuf(srcs[i])):
                errs.append(D.geterror())

    th = [threading.Thread(target=work, args=(i,)) for i in range(6)]
    for t in th:
        t.start()
    for t in th:
        t.join()
    assert not errs, errs
    for i in range(6):
        ref = np.zeros_like(outs[i])
        oracle.resize(oracle.Img(ref), oracle.Img(srcs[i]))
        check(outs[i], ref)


@pytest.mark.parametrize("dt,st", [(np.uint8, np.uint16), (np.uint8, np.float16),
                                   (np.float16, np.uint8), (np.float16, np.uint16),
                                   (np.uint16, np.uint8), (np.uint16, np.float16)])
def test_non_native_type_pairs_go_through_float(oiio, oracle, dt, st):
    """OIIO_DISPATCH_COMMON_TYPES2 computes the six non-native (dst,src)
    pairs into a float temporary and converts back (imagebufalgo_util.h:
    463-478, 540-546); here the temporary lives on the device."""
    src = synth.image(150, 110, 3, st, seed=301)
    tmp = np.zeros((50, 70, 3), np.float32)
    oracle.resize(oracle.Img(tmp), oracle.Img(src))
    ref = oracle.quantize(tmp, dt)
    out = np.zeros((50, 70, 3), dt)
    assert oiio.ImageBufAlgo.resize(oiio.ImageBuf(out), oiio.ImageBuf(src))
    check(out, ref)
    # resample, nearest: exact
    tmp = np.zeros((50, 70, 3), np.float32)
    oracle.resample(oracle.Img(tmp), oracle.Img(src), False)
    out = np.zeros((50, 70, 3), dt)
    assert oiio.ImageBufAlgo.resample(oiio.ImageBuf(out), oiio.ImageBuf(src),
                                      interpolate=False)
    check(out, oracle.quantize(tmp, dt), exact=True)


# ---- rangecompress / rangeexpand and highlight compensation ---------------
def _range_both(oiio, oracle, fn, src, dst_dtype, useluma, alpha, roi=None,
                chrange=None, prefill=None):
    h, w, c = src.shape
    g = np.zeros((h, w, c), dst_dtype) if prefill is None else prefill.copy()
    o = g.copy()
    getattr(oracle, fn)(oracle.Img(o), oracle.Img(src), useluma=useluma,
                        alpha_channel=alpha, roi=roi, chrange=chrange)
    S, D = oiio.ImageBuf(src), oiio.ImageBuf(g)
    S.spec().alpha_channel = alpha
    r = None
    if roi or chrange:
        x0, x1, y0, y1 = roi if roi else (0, w, 0, h)
        c0, c1 = chrange if chrange else (0, c)
        r = oiio.ROI(x0, x1, y0, y1, 0, 1, c0, c1)
    ok = getattr(oiio.ImageBufAlgo, fn)(D, S, useluma=useluma, roi=r)
    assert ok, D.geterror()
    return g, o


@pytest.mark.parametrize("fn", ["rangecompress", "rangeexpand"])
@pytest.mark.parametrize("dt,st", [("float", "float"), ("uint8", "uint8"),
                                   ("half", "float"), ("uint16", "half"),
                                   ("float", "uint16")])
def test_range_kernels(oiio, oracle, fn, dt, st):
    """logf / expf are the device's: float results to 2e-6 of the value range,
    integers +-1 LSB."""
    for nch, alpha, luma in [(4, 3, False), (4, 3, True), (3, -1, True),
                             (3, 1, True), (1, -1, False), (5, 4, False)]:
        src = synth.image(67, 41, nch, _np_dtype(st), seed=90 + nch)
        if st == "float":
            src = (src * np.float32(8.0) - np.float32(2.0)).astype(np.float32)
            if fn == "rangeexpand":
                src = (src * np.float32(0.25)).astype(np.float32)
        g, o = _range_both(oiio, oracle, fn, src, _np_dtype(dt), luma, alpha)
        if g.dtype in (np.uint8, np.uint16):
            assert np.abs(g.astype(int) - o.astype(int)).max() <= 1
        else:
            fin = np.isfinite(o.astype(np.float64))
            assert np.array_equal(fin, np.isfinite(g.astype(np.float64)))
            scale = max(1.0, np.abs(o[fin]).astype(np.float64).max())
            tol = 2e-6 if g.dtype == np.float32 else 2.0 ** -10
            assert np.abs(g[fin].astype(np.float64) - o[fin]).max() <= tol * scale


def test_range_roi_channels_inplace_and_goldens(oiio, oracle):
    src = (synth.image(50, 30, 4, np.float32, seed=95) * np.float32(5)).astype(np.float32)
    pre = synth.image(50, 30, 4, np.float32, seed=96)
    g, o = _range_both(oiio, oracle, "rangecompress", src, np.float32, False, 3,
                       roi=(5, 44, 3, 21), chrange=(1, 3), prefill=pre)
    assert np.abs(g - o).max() <= 2e-6 * 5
    assert np.array_equal(g[:3], pre[:3]) and np.array_equal(g[..., 0], pre[..., 0])
    # in place on a device tensor: alpha untouched
    import torch
    t = torch.from_numpy(src).cuda()
    T = oiio.ImageBuf(t)
    assert oiio.ImageBufAlgo.rangecompress(T, T)
    o2 = src.copy()
    oracle.rangecompress(oracle.Img(o2), oracle.Img(o2), alpha_channel=3)
    g2 = t.cpu().numpy()
    assert np.array_equal(g2[..., 3], src[..., 3])
    assert np.abs(g2 - o2).max() <= 2e-6 * 5
    # the reference's known-answer images (oiiotool reads uint8 as float)
    import os
    G = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden",
                             "pixelmath_ref.npz"))
    f = G["tahoe_small"].astype(np.float32) * np.float32(1 / 255.0)
    for luma, sfx in ((False, ""), (True, "_luma")):
        d = np.zeros(f.shape, np.uint8)
        D, S = oiio.ImageBuf(d), oiio.ImageBuf(f)
        assert oiio.ImageBufAlgo.rangecompress(D, S, useluma=luma)
        assert np.abs(d.astype(int) - G["rangecompress" + sfx].astype(int)).max() <= 1
        assert (d != G["rangecompress" + sfx]).mean() < 1e-3
        f2 = G["rangecompress" + sfx].astype(np.float32) * np.float32(1 / 255.0)
        d2 = np.zeros(f.shape, np.uint8)
        assert oiio.ImageBufAlgo.rangeexpand(oiio.ImageBuf(d2), oiio.ImageBuf(f2),
                                             useluma=luma)
        assert np.abs(d2.astype(int) - G["rangeexpand" + sfx].astype(int)).max() <= 1
        assert (d2 != G["rangeexpand" + sfx]).mean() < 1e-3


@pytest.mark.parametrize("dt,st,nch", [("float", "float", 4), ("half", "half", 4),
                                       ("uint8", "uint8", 3), ("float", "float", 3)])
def test_resize_highlightcomp(oiio, oracle, dt, st, nch):
    """oiiotool --resize:highlightcomp=1 (oiiotool.cpp:4644-4663): compress into
    a temporary of the source type, resize, expand the result in place."""
    src = synth.image(120, 90, nch, _np_dtype(st), seed=101)
    if st in ("float", "half"):
        src = (src.astype(np.float32) ** 4 * 40).astype(_np_dtype(st))  # HDR highlights
    alpha = 3 if nch == 4 else -1
    for shape in [(45, 60), (200, 260)]:
        tmp = src.copy()
        oracle.rangecompress(oracle.Img(tmp), oracle.Img(src), alpha_channel=alpha)
        o = np.zeros(shape + (nch,), _np_dtype(dt))
        oracle.resize(oracle.Img(o), oracle.Img(tmp), "lanczos3")
        oracle.rangeexpand(oracle.Img(o), oracle.Img(o), alpha_channel=alpha)
        g = np.zeros_like(o)
        D, S = oiio.ImageBuf(g), oiio.ImageBuf(src)
        S.spec().alpha_channel = alpha
        assert oiio.ImageBufAlgo.resize(D, S, filtername="lanczos3",
                                        highlightcomp=True), D.geterror()
        if g.dtype == np.uint8:
            # the resize's +-1 LSB sits under rangeexpand, whose slope reaches
            # x/b = 5.4 at the top of the range: a handful of pixels move by up
            # to 6 LSB there, the rest agree
            d = np.abs(g.astype(int) - o.astype(int))
            assert d.max() <= 6 and (d > 0).mean() < 2e-3
        else:
            # expf amplifies the resize's 1e-6 by 1/b = 5.4: 5e-5 of the range
            scale = max(1.0, float(np.abs(o.astype(np.float64)).max()))
            tol = 5e-5 if g.dtype == np.float32 else 2.0 ** -9
            assert np.abs(g.astype(np.float64) - o.astype(np.float64)).max() <= tol * scale
        # and it differs from the plain resize (the option is not ignored)
        p = np.zeros_like(o)
        assert oiio.ImageBufAlgo.resize(oiio.ImageBuf(p), S, filtername="lanczos3")
        if st != "uint8":
            assert np.abs(p.astype(np.float64) - g.astype(np.float64)).max() > 1e-3


@pytest.mark.gpu
@pytest.mark.parametrize("st,dt", [("float", "float"), ("half", "half"), ("half", "float"),
                                   ("float", "half")])
def test_resize_highlightcomp_fused(oiio, oracle, st, dt):
    """Device-resident float / half images of a shape the wide stream kernel takes:
    highlight compensation runs INSIDE the resize kernel (one launch) -- compress as
    the V pass loads, expand as the H pass stores, with the roundings the three
    passes make through their half temporaries -- and matches the oracle's
    compress -> resize -> expand."""
    import torch
    w, h = 3072, 2304  # 6 strips x 36 bands: enough items for the wide geometry
    src = (synth.image(w, h, 4, np.float32, seed=131) ** 4 * np.float32(20)).astype(_np_dtype(st))
    tmp = src.copy()
    oracle.rangecompress(oracle.Img(tmp), oracle.Img(src), alpha_channel=3)
    ref = np.zeros((h // 2, w // 2, 4), _np_dtype(dt))
    oracle.resize(oracle.Img(ref), oracle.Img(tmp), "lanczos3")
    oracle.rangeexpand(oracle.Img(ref), oracle.Img(ref), alpha_channel=3)
    tdt = {"float": torch.float32, "half": torch.float16}
    S = oiio.ImageBuf(torch.from_numpy(src).cuda())
    dst = torch.zeros((h // 2, w // 2, 4), dtype=tdt[dt], device="cuda")
    D = oiio.ImageBuf(dst)
    S.spec().alpha_channel = D.spec().alpha_channel = 3
    oiio.ImageBufAlgo.force_report = True
    try:
        assert oiio.ImageBufAlgo.resize(D, S, filtername="lanczos3", highlightcomp=True), D.geterror()
    finally:
        oiio.ImageBufAlgo.force_report = False
    assert oiio.ImageBufAlgo.last_report.kernels_launched == 1  # fused, not three passes
    got = dst.cpu().numpy()
    scale = max(1.0, float(np.abs(ref.astype(np.float64)).max()))
    tol = 5e-5 if dt == "float" and st == "float" else 2.0 ** -9
    err = np.abs(got.astype(np.float64) - ref.astype(np.float64)).max()
    assert err <= tol * scale, (err, scale)


def test_mip_chain_highlightcomp(oiio, oracle):
    lvl0 = (synth.image(96, 64, 4, np.float32, seed=111) ** 4 * np.float32(30)).astype(np.float32)
    ref = oracle.mip_chain(lvl0, "lanczos3", highlightcomp=True, alpha_channel=3)
    chain = oiio.MipChain(lvl0, "lanczos3", highlightcomp=True)
    assert chain.nlevels == len(ref) + 1
    # level 0 is left intact on the device
    assert np.array_equal(chain.download(0), lvl0)
    for i, r in enumerate(ref):
        g = chain.download(i + 1)
        scale = max(1.0, float(np.abs(r).max()))
        assert np.abs(g - r).max() <= 5e-5 * (i + 1) * scale, "level %d" % (i + 1)
    # box chains ignore the option, as maketx does
    a = oiio.MipChain(lvl0, "box", highlightcomp=True).download(1)
    b = oiio.MipChain(lvl0, "box").download(1)
    assert np.array_equal(a, b)


@pytest.mark.parametrize("st,dt", [("float", "float"), ("uint16", "uint16")])
def test_host_pipeline_pageable_and_pinned_equal_device(oiio, oracle, st, dt):
    """Images above 32 MB take the 3-stream row-band pipeline; pageable host
    memory (a numpy array, what an ImageBuf normally owns) additionally goes
    through the pinned bounce rings.  Pageable, pinned and device-resident
    calls must give the same bytes, and match the oracle on a band."""
    import torch
    w, h, c = 3000, 2300, 4
    src = synth.image(w, h, c, _np_dtype(st), seed=171)
    dw, dh = 1500, 1100
    IBA = oiio.ImageBufAlgo

    def tt(a):
        return (torch.from_numpy(a.view(np.int16)).view(torch.uint16)
                if a.dtype == np.uint16 else torch.from_numpy(a))
    out_page = np.zeros((dh, dw, c), _np_dtype(dt))
    assert IBA.resize(oiio.ImageBuf(out_page), oiio.ImageBuf(src), filtername="lanczos3")
    ps = tt(src).pin_memory()
    pd = tt(np.zeros_like(out_page)).pin_memory()
    assert IBA.resize(oiio.ImageBuf(pd), oiio.ImageBuf(ps), filtername="lanczos3")
    dd = tt(np.zeros_like(out_page)).cuda()
    assert IBA.resize(oiio.ImageBuf(dd), oiio.ImageBuf(tt(src).cuda()),
                      filtername="lanczos3")
    b = lambda t: t.cpu().view(torch.uint8).numpy()
    assert np.array_equal(out_page.view(np.uint8), b(pd))
    assert np.array_equal(out_page.view(np.uint8), b(dd))
    ref = np.zeros_like(out_page)
    oracle.resize(oracle.Img(ref), oracle.Img(src), "lanczos3", roi=(0, dw, 500, 532))
    check(out_page[500:532], ref[500:532])


def test_nonfinite_redo_is_local(oiio, oracle):
    """One Inf in a large image: only the dirty-map cells around it are
    re-done (bit-exact there, NaN/Inf pattern as the reference), the rest of
    the image keeps the fused kernel's result; and the redo costs a fraction
    of a full exact pass."""
    import torch
    src = synth.image(2048, 1536, 4, np.float32, seed=181)
    src[700, 1000, 2] = np.inf
    src[701, 1000, 0] = np.nan
    t = torch.from_numpy(src).cuda()
    out = torch.zeros((768, 1024, 4), dtype=torch.float32, device="cuda")
    IBA = oiio.ImageBufAlgo
    assert IBA.resize(oiio.ImageBuf(out), oiio.ImageBuf(t), filtername="lanczos3")
    got = out.cpu().numpy()
    ref = np.zeros_like(got)
    y0, y1 = 320, 384           # the band that holds the defect (350 = 700 / 2)
    oracle.resize(oracle.Img(ref), oracle.Img(src), "lanczos3", roi=(0, 1024, y0, y1))
    check(got[y0:y1], ref[y0:y1])
    rep = oiio.ImageBufAlgo.last_report
    if rep is not None and rep.nonfinite_redo == 1:
        # the cell holding the defect was re-done by the exact kernel
        assert np.array_equal(got[336:352, 384:512], ref[336:352, 384:512], equal_nan=True)
    assert np.isnan(got[350, 500]).any() or np.isinf(got[350, 500]).any()
    # timing: redo of a few cells vs the whole image through the exact kernel
    def timed(fn, n=5):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        fn()
        torch.cuda.synchronize()
        e0.record()
        for _ in range(n):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / n
    D, S = oiio.ImageBuf(out), oiio.ImageBuf(t)
    local = timed(lambda: IBA.resize(D, S, filtername="lanczos3"))
    full = timed(lambda: IBA.resize(D, S, filtername="lanczos3", path=oiio.PATH_EXACT))
    assert local < 0.6 * full, (local, full)


@pytest.mark.parametrize("st,dt,nch", [("float", "float", 4), ("float", "float", 3),
                                       ("half", "half", 4), ("half", "float", 1)])
@pytest.mark.parametrize("filt,shape", [("lanczos3", (150, 200)), ("box", (150, 200)),
                                        ("mitchell", (100, 131)), ("triangle", (75, 100))])
def test_nonfinite_dense(oiio, oracle, st, dt, nch, filt, shape):
    """NaN / +Inf / -Inf sprinkled over 1 % of the source values: the output's
    non-finite pattern is the reference's (zero weights never touch them,
    imagebufalgo_xform.cpp:749-759) and every finite output is within tolerance."""
    rng = np.random.RandomState(5)
    src = synth.image(400, 300, nch, np.float32, seed=91)
    hit = rng.rand(*src.shape) < 0.01
    src[hit] = rng.choice(np.array([np.nan, np.inf, -np.inf], np.float32), size=int(hit.sum()))
    src = src.astype(_np_dtype(st))
    g, o, rep = run_both(oiio, oracle, src, shape, _np_dtype(dt), filt)
    check(g, o)
    assert rep.path_used == oiio.PATH_FUSED
    assert np.isfinite(o.astype(np.float64)).any() and not np.isfinite(o.astype(np.float64)).all()


@pytest.mark.parametrize("nch", [5, 6, 9])
@pytest.mark.parametrize("st,dt", [("float", "float"), ("uint8", "uint8"), ("half", "float")])
def test_more_than_four_channels_split_into_groups(oiio, oracle, nch, st, dt):
    """Multi-layer images: channel groups of <= 4 go through the fused kernel
    (the report says so) and are put back; ROI and prefill respected; host and
    device-resident."""
    import torch
    src = synth.image(130, 90, nch, _np_dtype(st), seed=191)
    pre = synth.image(65, 45, nch, _np_dtype(dt), seed=192)
    g, o, rep = run_both(oiio, oracle, src, (45, 65), _np_dtype(dt), "lanczos3",
                         roi=(3, 60, 2, 40), prefill=pre)
    check(g, o)
    assert rep.path_used == oiio.PATH_FUSED
    if dt != "uint16" and st != "uint16":
        def up(a):
            return torch.from_numpy(a).cuda()
        D = oiio.ImageBuf(up(pre))
        assert oiio.ImageBufAlgo.resize(D, oiio.ImageBuf(up(src)), filtername="lanczos3",
                                        roi=oiio.ROI(3, 60, 2, 40, 0, 1, 0, nch))
        assert np.array_equal(D.get_pixels().view(np.uint8), g.view(np.uint8))


def test_source_with_more_channels_than_destination(oiio, oracle):
    """resize_ reads the first dst.nchannels channels of a wider source; the
    channel-group copy drops the rest and the fused kernel runs."""
    src = synth.image(120, 80, 4, np.float32, seed=195)
    ref = np.zeros((40, 60, 3), np.float32)
    oracle.resize(oracle.Img(ref), oracle.Img(np.ascontiguousarray(src[..., :3])), "lanczos3")
    got = np.zeros_like(ref)
    D = oiio.ImageBuf(got)
    assert oiio.ImageBufAlgo.resize(D, oiio.ImageBuf(src), filtername="lanczos3"), D.geterror()
    check(got, ref)
    assert oiio.ImageBufAlgo.last_report.path_used == oiio.PATH_FUSED


@pytest.mark.gpu
@pytest.mark.parametrize("st", ["float", "uint8", "half"])
def test_device_views_with_gaps_between_pixels(oiio, oracle, st):
    """Device images whose pixels are not contiguous in x (channel sub-views of
    wider buffers, on either side) are packed by the channel-group copy and run
    the fused path instead of the single-pass exact kernel; the bytes between
    the destination's pixels are left alone."""
    import torch
    tdt = {"float": torch.float32, "uint8": torch.uint8, "half": torch.float16}[st]
    wide = synth.image(300, 200, 8, _np_dtype(st), seed=410)       # 8 channels, use 1..4
    src_view = torch.from_numpy(wide).cuda()[..., 1:5]
    assert not src_view.is_contiguous()
    canvas = torch.full((100, 150, 6), 7, dtype=tdt, device="cuda")  # 6 channels, write 2..5
    dst_view = canvas[..., 2:6]
    D, S = oiio.ImageBuf(dst_view), oiio.ImageBuf(src_view)
    oiio.ImageBufAlgo.force_report = True
    try:
        assert oiio.ImageBufAlgo.resize(D, S, filtername="lanczos3"), D.geterror()
    finally:
        oiio.ImageBufAlgo.force_report = False
    torch.cuda.synchronize()
    rep = oiio.ImageBufAlgo.last_report
    assert rep.path_used == oiio.PATH_FUSED and rep.kernels_launched >= 3  # pack, resize, unpack
    ref = np.zeros((100, 150, 4), _np_dtype(st))
    oracle.resize(oracle.Img(ref), oracle.Img(np.ascontiguousarray(wide[..., 1:5])), "lanczos3")
    got = canvas.cpu().numpy()
    check(np.ascontiguousarray(got[..., 2:6]), ref)
    assert np.all(got[..., :2] == 7)     # the gap channels are untouched


@pytest.mark.parametrize("st", ["float", "uint8"])
def test_overscan_interior_fused_frame_exact(oiio, oracle, st):
    """A source whose data window extends beyond its display window: device
    resident, the interior (all taps inside the data window) goes through the
    fused kernel, the frame around it through the exact kernel; the whole
    image must match the oracle, seams included."""
    import torch
    sw, sh = 700, 500
    src = synth.image(sw, sh, 4, _np_dtype(st), seed=201)
    swin = (-4, -3, (0, 0, sw - 8, sh - 6))              # 4 / 3 px of overscan all round
    dw, dh = 326, 234
    dwin = (0, 0, (0, 0, dw, dh))
    ref = np.zeros((dh, dw, 4), _np_dtype(st))
    oracle.resize(oracle.Img(ref, *dwin), oracle.Img(src, *swin), "lanczos3")
    t = torch.from_numpy(src).cuda()
    out = torch.zeros((dh, dw, 4), dtype=t.dtype, device="cuda")
    S, D = oiio.ImageBuf(t), oiio.ImageBuf(out)
    for B, (x, y, full) in ((S, swin), (D, dwin)):
        sp = B.spec()
        sp.x, sp.y = x, y
        sp.full_x, sp.full_y, sp.full_width, sp.full_height = full
    assert oiio.ImageBufAlgo.resize(D, S, filtername="lanczos3"), D.geterror()
    got = out.cpu().numpy()
    check(got, ref)
    # the same through the exact kernel only: the interior differs from it in
    # the last bits (fused), so the fused kernel did run there
    out2 = torch.zeros_like(out)
    D2 = oiio.ImageBuf(out2)
    sp = D2.spec()
    sp.full_x, sp.full_y, sp.full_width, sp.full_height = dwin[2]
    assert oiio.ImageBufAlgo.resize(D2, S, filtername="lanczos3", path=oiio.PATH_EXACT)
    g2 = out2.cpu().numpy()
    check(g2, ref)
    # the first row / column need taps beyond the data window: exact kernel
    assert np.array_equal(got[:1], g2[:1]) and np.array_equal(got[:, :1], g2[:, :1])
    if st == "float":
        assert not np.array_equal(got[60:160, 80:240], g2[60:160, 80:240])


# ---- MIP chain of a level 0 that is offset / cropped (orig_was_overscan) ----
@pytest.mark.gpu
@pytest.mark.parametrize("filt", ["box", "lanczos3"])
@pytest.mark.parametrize("origin,full", [((37, 11), None), ((0, 0), (0, 0, 200, 150)),
                                         ((-6, -4), (0, 0, 180, 120))])
def test_mip_chain_offset_level0(oiio, oracle, filt, origin, full):
    """write_mipmap sends an image whose windows are not plain (non-zero origin,
    crop, overscan: orig_was_overscan, maketexture.cpp:700-703) through
    resize() with the named filter even for "box" (:881), level 0 keeping its
    origin with display == data (:876-877); levels >= 1 sit at the origin."""
    w, h = 192, 128
    lvl0 = synth.image(w, h, 4, np.float32, seed=77)
    spec = oiio.ImageSpec(w, h, 4, "float")
    spec.x, spec.y = origin
    if full is not None:
        spec.full_x, spec.full_y, spec.full_width, spec.full_height = full
    else:
        spec.full_x, spec.full_y, spec.full_width, spec.full_height = origin[0], origin[1], w, h
    chain = oiio.MipChain(oiio.ImageBuf(lvl0, spec), filt)
    # the oracle's loop, composed by hand with the reference's windows
    big = oracle.Img(lvl0, x=origin[0], y=origin[1], full=(origin[0], origin[1], w, h))
    lv = 1
    while big.arr.shape[1] > 1 or big.arr.shape[0] > 1:
        bh, bw, _ = big.arr.shape
        small = np.empty((max(1, bh // 2), max(1, bw // 2), 4), np.float32)
        oracle.resize(oracle.Img(small), big, filt)
        got = chain.download(lv)
        assert got.shape == small.shape
        assert np.max(np.abs(got - small)) <= 1e-5 * lv, (lv, np.max(np.abs(got - small)))
        big = oracle.Img(small)
        lv += 1
    assert lv == chain.nlevels


# ---- thumbnail-scale minification: windows wider than any fused instantiation ----
def _separable_truth(oiio, src, dst_shape, filt):
    """resize_ (imagebufalgo_xform.cpp:595-826) for plain windows, evaluated in
    float64: the reference's per-axis taps (float32 positions and weights, as it
    computes them), normalised per axis, applied as Wy @ src @ Wx^T in double."""
    import math
    dh, dw = dst_shape
    sh, sw, _ = src.shape
    native = {"lanczos3": 6.0, "triangle": 2.0, "mitchell": 4.0, "blackman-harris": 3.0}[filt]
    F = oiio.Filter2D.create(filt, native * max(1.0, dw / sw), native * max(1.0, dh / sh))

    def axis(n_dst, n_src, f1, width):
        ratio = np.float32(n_dst) / np.float32(n_src)
        rad = int(math.ceil((width / 2.0) / float(ratio)))
        W = np.zeros((n_dst, n_src), np.float64)
        for d in range(n_dst):
            s = np.float32(d + 0.5) * (np.float32(1.0) / np.float32(n_dst))
            sf = np.float32(s * np.float32(n_src))
            si = int(math.floor(sf))
            frac = np.float32(sf - np.float32(si))
            w = np.array([f1(float(ratio * np.float32(i - rad - (frac - np.float32(0.5)))))
                          for i in range(2 * rad + 1)], np.float32)
            tot = np.float32(w.sum(dtype=np.float32))
            if tot != 0:
                w = w / tot
            for i, wi in enumerate(w):
                W[d, min(max(si - rad + i, 0), n_src - 1)] += float(wi)
        return W
    Wy = axis(dh, sh, F.yfilt, native * max(1.0, dh / sh))
    Wx = axis(dw, sw, F.xfilt, native * max(1.0, dw / sw))
    return np.einsum("yr,rxc,dx->ydc", Wy, src.astype(np.float64), Wx)



@pytest.mark.gpu
@pytest.mark.parametrize("st,dt,nch", [("float", "float", 4), ("uint8", "uint8", 3),
                                       ("half", "half", 4), ("uint16", "uint16", 1),
                                       ("float", "uint8", 3), ("float", "float", 5)])
@pytest.mark.parametrize("filt", ["lanczos3", "triangle", "mitchell"])
def test_thumbnail_minification_wide_windows(oiio, oracle, st, dt, nch, filt):
    """1200 x 900 -> 23 x 17 (52:1; lanczos3: 313 taps per axis): no fused kernel
    holds such a window; the two-pass wide kernels (wide_kernel.cu) run instead
    of the O(taps^2) exact kernel and match the oracle like the fused ones do.
    Host and device-resident, with a partial ROI and prefilled destination."""
    src = synth.image(1200, 900, nch, _np_dtype(st), seed=700)
    pre = synth.image(23, 17, nch, _np_dtype(dt), seed=701)
    g, o, rep = run_both(oiio, oracle, src, (17, 23), _np_dtype(dt), filt)
    if dt == "float":
        # At this scale the reference adds ~98 000 float products per pixel one
        # after the other: its own rounding noise (~ sqrt(taps) * 2^-24) is above
        # 1e-5, so "within 1e-5 of the reference" is not defined by the arithmetic
        # any more.  Pin both to the float64 evaluation of the same taps instead:
        # the two-pass result must be at least as close to it as the reference's,
        # and within the reference's noise of the reference.
        truth = _separable_truth(oiio, src, (17, 23), filt)
        e_ref = np.abs(o.astype(np.float64) - truth).max()
        e_gpu = np.abs(g.astype(np.float64) - truth).max()
        assert e_gpu <= max(e_ref, 2e-6), (e_gpu, e_ref)
        assert np.abs(g.astype(np.float64) - o.astype(np.float64)).max() <= 2 * max(e_ref, 5e-6)
    else:
        check(g, o)
    assert rep.path_used == oiio.PATH_FUSED
    if nch <= 4:  # (five channels: two channel groups, each through the wide kernels)
        assert rep.kernels_launched == 2
    g, o, rep = run_both(oiio, oracle, src, (17, 23), _np_dtype(dt), filt, roi=(3, 20, 2, 15),
                         prefill=pre)
    if dt == "float":
        assert np.abs(g.astype(np.float64) - o.astype(np.float64)).max() <= 5e-5
    else:
        check(g, o)
    assert np.array_equal(g[:2], pre[:2]) and np.array_equal(g[:, :3], pre[:, :3])


@pytest.mark.gpu
def test_thumbnail_cropped_source_and_one_axis(oiio, oracle):
    """A cropped source window (black beyond the data window, clamped taps merged)
    and minification along one axis only."""
    src = synth.image(1000, 64, 4, np.float32, seed=702)
    g, o, rep = run_both(oiio, oracle, src, (64, 20), np.float32, "lanczos3",
                         src_win=(7, 0, (0, 0, 1020, 64)))
    check(g, o)
    assert rep.kernels_launched == 2
    src = synth.image(64, 1000, 3, np.uint16, seed=703)
    g, o, rep = run_both(oiio, oracle, src, (20, 64), np.uint16, "blackman-harris")
    check(g, o)


# ---- b2r_resize_batch: many frames of one shape in one launch -------------
@pytest.mark.gpu
@pytest.mark.parametrize("st,dt,nch,filt,sshape,dshape,nfr", [
    ("uint16", "uint16", 3, "lanczos3", (432, 768), (216, 384), 5),
    ("float", "float", 4, "lanczos3", (600, 1100), (300, 550), 5),
    ("half", "half", 4, "blackman-harris", (400, 640), (200, 320), 5),
    ("float", "float", 3, "mitchell", (300, 500), (170, 260), 5),
    ("uint8", "uint8", 3, "mitchell", (90, 160), (360, 640), 5),      # expand kernel: frame by frame
    # enough frames that the schedule search gives every frame ONE band (no halo rows)
    ("uint16", "uint16", 3, "lanczos3", (432, 768), (216, 384), 80),
    ("uint8", "uint8", 4, "lanczos3", (432, 768), (216, 384), 150),
    ("half", "float", 4, "lanczos3", (432, 1100), (216, 550), 37),
])
def test_resize_batch_equals_single_frames(oiio, st, dt, nch, filt, sshape, dshape, nfr):
    """Frames of one shape through ONE persistent launch give exactly what
    one call per frame gives (and the shapes the stream kernel does not take
    fall back to frame-by-frame).  The band count of the batch plan differs
    from the single-frame plan's (fused_plan.cpp searches it per batch size):
    the results must not."""
    import torch
    IBA = oiio.ImageBufAlgo

    def T(a):
        t = torch.from_numpy(a.view(np.int16) if a.dtype == np.uint16 else a)
        return (t.view(torch.uint16) if a.dtype == np.uint16 else t).cuda()
    base = [T(synth.image(sshape[1], sshape[0], nch, _np_dtype(st), seed=300 + i))
            for i in range(min(nfr, 5))]
    def rolled(t, k):  # a different picture per frame (uint16 rolls as bytes)
        return torch.roll(t.view(torch.uint8), k, 0).contiguous().view(t.dtype)
    srcs = [base[i] if i < len(base) else rolled(base[i % len(base)], 7 * i) for i in range(nfr)]
    tdt = {"uint8": torch.uint8, "uint16": torch.uint16, "half": torch.float16,
           "float": torch.float32}[dt]
    one = [torch.zeros((dshape[0], dshape[1], nch), dtype=tdt, device="cuda") for _ in range(nfr)]
    many = [torch.zeros_like(t) for t in one]
    for s_, d in zip(srcs, one):
        assert IBA.resize(oiio.ImageBuf(d), oiio.ImageBuf(s_), filtername=filt)
    assert IBA.resize_batch([oiio.ImageBuf(d) for d in many], [oiio.ImageBuf(s_) for s_ in srcs],
                            filtername=filt, report=True)
    torch.cuda.synchronize()
    for a, b in zip(one, many):
        assert torch.equal(a.view(torch.uint8), b.view(torch.uint8))
    if st != "uint8":
        assert IBA.last_report.kernels_launched == 1  # one launch for all the frames
