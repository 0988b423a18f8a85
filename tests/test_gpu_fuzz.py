"""Seeded differential fuzz of the resize path: random sizes (incl. 1-pixel
axes and prime sizes), ratios from 8x down to 8x up (independently per axis),
every filter, every native type pair, 1-5 channels, cropped /
overscanned source windows, offset destination windows, partial ROIs, random
filter widths -- CUDA path through the C ABI against the CPU oracle."""
import numpy as np
import pytest

import synth
from test_gpu_parity import run_both, check, _np_dtype

pytestmark = pytest.mark.gpu

# B2R_FUZZ_SEED=<n> shifts every seed: fresh random cases for an extra sweep
import os
SEED = int(os.environ.get("B2R_FUZZ_SEED", "0")) * 100003

FILTERS = ["", "box", "triangle", "gaussian", "sharp-gaussian", "catmull-rom",
           "blackman-harris", "sinc", "lanczos3", "radial-lanczos3", "nuke-lanczos6",
           "mitchell", "bspline", "disk", "cubic", "keys", "simon", "rifman"]
TYPES = ["float", "half", "uint16", "uint8"]


def _case(rng):
    sw = int(rng.choice([1, 2, 3, 5, 17, 31, 64, 97, 128, 200, 257, 400]))
    sh = int(rng.choice([1, 2, 3, 7, 16, 33, 64, 90, 131, 256, 300]))
    rx = float(rng.choice([0.125, 0.25, 0.37, 0.5, 0.75, 1.0, 1.3, 2.0, 3.0, 4.0, 8.0]))
    ry = float(rng.choice([0.125, 0.25, 0.4, 0.5, 0.8, 1.0, 1.5, 2.0, 2.5, 4.0, 8.0]))
    dw = max(1, min(700, int(round(sw * rx))))
    dh = max(1, min(700, int(round(sh * ry))))
    nch = int(rng.choice([1, 2, 3, 4, 5]))
    # the ten native (dst, src) pairs (the others are covered by
    # test_non_native_type_pairs_go_through_float)
    dt = str(rng.choice(TYPES))
    st = str(rng.choice(TYPES)) if dt == "float" else str(rng.choice([dt, "float"]))
    filt = str(rng.choice(FILTERS))
    fw = float(rng.choice([0.0, 0.0, 0.0, 2.5, 5.0]))
    src_win = dst_win = roi = None
    if rng.rand() < 0.35:   # source data window != display window (crop or overscan)
        ox, oy = int(rng.randint(-6, 7)), int(rng.randint(-6, 7))
        src_win = (ox, oy, (0, 0, max(1, sw + int(rng.randint(-5, 6))),
                            max(1, sh + int(rng.randint(-5, 6)))))
    if rng.rand() < 0.35:   # destination data window inside / around its display window
        ox, oy = int(rng.randint(-4, 5)), int(rng.randint(-4, 5))
        dst_win = (ox, oy, (0, 0, max(1, dw + int(rng.randint(-3, 4))),
                            max(1, dh + int(rng.randint(-3, 4)))))
    if rng.rand() < 0.3:
        dx, dy = (dst_win[0], dst_win[1]) if dst_win else (0, 0)
        x0 = dx + int(rng.randint(0, max(1, dw // 2)))
        y0 = dy + int(rng.randint(0, max(1, dh // 2)))
        roi = (x0, min(dx + dw, x0 + 1 + int(rng.randint(0, dw))),
               y0, min(dy + dh, y0 + 1 + int(rng.randint(0, dh))))
    # radial-lanczos3 with a narrow FIXED width normalises by a weight sum that
    # can nearly cancel: the result then hangs on the last bit of sinf (device
    # vs glibc) -- in the reference itself as much as here.  Keep its own width.
    if filt == "radial-lanczos3":
        fw = 0.0
    return dict(sw=sw, sh=sh, dw=dw, dh=dh, nch=nch, st=str(st), dt=str(dt), filt=filt,
                fw=fw, src_win=src_win, dst_win=dst_win, roi=roi)


NATIVE_WIDTH = {"box": 1, "triangle": 2, "gaussian": 3, "sharp-gaussian": 2, "catmull-rom": 4,
                "blackman-harris": 3, "sinc": 4, "lanczos3": 6, "nuke-lanczos6": 6, "mitchell": 4,
                "bspline": 4, "cubic": 4, "keys": 4, "simon": 4, "rifman": 4}


def _axis_amplification(f1, width, dfx, dfw, sfx, sfw, x0, x1):
    """max over the dst range of sum|w| / |sum w| for the reference's taps
    (imagebufalgo_xform.cpp:626-660): what normalising by the weight sum
    multiplies a rounding error by."""
    import math
    ratio = float(np.float32(dfw) / np.float32(sfw))
    rad = int(math.ceil((width / 2.0) / ratio))
    worst = 1.0
    for x in range(x0, x1):
        sf = sfx + ((x - dfx + 0.5) / dfw) * sfw
        frac = sf - math.floor(sf)
        w = np.array([f1(ratio * (i - rad - (frac - 0.5))) for i in range(2 * rad + 1)])
        tot = w.sum()
        if tot != 0.0:
            worst = max(worst, float(np.abs(w).sum() / abs(tot)))
    return worst


def _int_lsb_tol(oiio, oracle, c, src):
    """LSB tolerance of an integer destination: 1 for a well-conditioned filter.
    The reference normalises each axis by its weight sum; for some filters and
    widths (sinc at 4x upsizing: 149; rifman with a fixed width of 2.5 at 1.3x:
    thousands) that sum nearly cancels and the normalised weights are huge -- in
    the reference itself, whose float result then carries eps x sum|w| of
    rounding noise.  An integer output inside the clamp range shows it in LSBs;
    the tolerance grows with the amplification of the case's own taps (the float
    comparison scales with the output magnitude already)."""
    if c["dt"] not in ("uint8", "uint16"):
        return 1
    sw, sh, dw, dh = c["sw"], c["sh"], c["dw"], c["dh"]
    sfx, sfy, sfw, sfh = c["src_win"][2] if c["src_win"] else (0, 0, sw, sh)
    dfx, dfy, dfw, dfh = c["dst_win"][2] if c["dst_win"] else (0, 0, dw, dh)
    dx0, dy0 = (c["dst_win"][0], c["dst_win"][1]) if c["dst_win"] else (0, 0)
    name = c["filt"] or ("blackman-harris" if (dfw > sfw or dfh > sfh) else "lanczos3")
    maxv = 255.0 if c["dt"] == "uint8" else 65535.0
    if name not in NATIVE_WIDTH:  # disk, radial-lanczos3: not separable; keep their own width
        return 1
    wx = c["fw"] if c["fw"] > 0 else NATIVE_WIDTH[name] * max(1.0, dfw / sfw)
    wy = c["fw"] if c["fw"] > 0 else NATIVE_WIDTH[name] * max(1.0, dfh / sfh)
    F = oiio.Filter2D.create(name, wx, wy)
    amp = (_axis_amplification(F.xfilt, wx, dfx, dfw, sfx, sfw, dx0, dx0 + dw)
           * _axis_amplification(F.yfilt, wy, dfy, dfh, sfy, sfh, dy0, dy0 + dh))
    return max(1, 1 + int(amp * maxv * 2.0 ** -22))


@pytest.mark.parametrize("chunk", range(16))
def test_resize_fuzz(oiio, oracle, chunk):
    rng = np.random.RandomState(20261020 + chunk + SEED)
    for k in range(60):
        c = _case(rng)
        src = synth.image(c["sw"], c["sh"], c["nch"], _np_dtype(c["st"]),
                          seed=1000 * chunk + k)
        pre = synth.image(c["dw"], c["dh"], c["nch"], _np_dtype(c["dt"]),
                          seed=5000 + 1000 * chunk + k)
        try:
            g, o, _ = run_both(oiio, oracle, src, (c["dh"], c["dw"]), _np_dtype(c["dt"]),
                               c["filt"], c["fw"], c["src_win"], c["dst_win"], c["roi"],
                               prefill=pre)
            check(g, o, lsb_tol=_int_lsb_tol(oiio, oracle, c, src))
        except AssertionError as e:
            raise AssertionError("case %d/%d %r: %s" % (chunk, k, c, e))


def test_regression_leading_rows_without_taps_host_source(oiio, oracle):
    """Found by the fuzz: a fixed filter width under 4x upsizing leaves the
    ROI's first dst rows with no contributing tap; the planner used to give
    them source row 0, outside the rows staged for a host source."""
    src = synth.image(257, 64, 2, np.float32, seed=1)
    pre = synth.image(95, 256, 2, np.float16, seed=2)
    g, o, rep = run_both(oiio, oracle, src, (256, 95), np.float16, "cubic", 2.5,
                         None, None, (23, 24, 20, 220), prefill=pre)
    check(g, o)


# ---------------------------------------------------------------- resample ----
@pytest.mark.parametrize("chunk", range(4))
def test_resample_fuzz(oiio, oracle, chunk):
    rng = np.random.RandomState(777 + chunk + SEED)
    for k in range(40):
        c = _case(rng)
        src = synth.image(c["sw"], c["sh"], c["nch"], _np_dtype(c["st"]), seed=k)
        interp = bool(rng.randint(0, 2))
        ref = synth.image(c["dw"], c["dh"], c["nch"], _np_dtype(c["dt"]), seed=99 + k)
        got = ref.copy()
        sx, sy, sfull = c["src_win"] if c["src_win"] else (0, 0, None)
        dx, dy, dfull = c["dst_win"] if c["dst_win"] else (0, 0, None)
        oracle.resample(oracle.Img(ref, dx, dy, dfull), oracle.Img(src, sx, sy, sfull),
                        interp, c["roi"])
        S, D = oiio.ImageBuf(src), oiio.ImageBuf(got)
        for B, (x, y, full), arr in ((S, (sx, sy, sfull), src), (D, (dx, dy, dfull), got)):
            sp = B.spec()
            sp.x, sp.y = x, y
            (sp.full_x, sp.full_y, sp.full_width, sp.full_height) = \
                full if full else (x, y, arr.shape[1], arr.shape[0])
        r = c["roi"]
        roi = oiio.ROI(r[0], r[1], r[2], r[3], 0, 1, 0, c["nch"]) if r else None
        assert oiio.ImageBufAlgo.resample(D, S, interpolate=interp, roi=roi), D.geterror()
        try:
            check(got, ref, exact=(got.dtype == np.float32))
        except AssertionError as e:
            raise AssertionError("case %d/%d %r interp=%s: %s" % (chunk, k, c, interp, e))


# -------------------------------------------------------------------- warp ----
@pytest.mark.parametrize("chunk", range(4))
def test_warp_fuzz(oiio, oracle, chunk):
    """Random affine / projective matrices (rotation, anisotropic scale down to
    1/6, shear, translation, mild perspective), every wrap mode, edgeclamp,
    cropped source windows, ROIs, 1-5 channels, the positive filters (the
    negative-lobe ones are ill-conditioned where a clipped footprint keeps
    few taps, see test_st_warp_parity)."""
    rng = np.random.RandomState(4242 + chunk + SEED)
    for k in range(25):
        sw, sh = int(rng.randint(8, 90)), int(rng.randint(8, 70))
        dw, dh = int(rng.randint(4, 80)), int(rng.randint(4, 60))
        nch = int(rng.choice([1, 2, 3, 4, 5]))
        st = str(rng.choice(TYPES))
        dt = str(rng.choice(TYPES))
        filt = str(rng.choice(["box", "triangle", "gaussian", "bspline", "disk",
                               "blackman-harris"]))
        while True:
            ang = rng.uniform(-3.14, 3.14)
            sx_, sy_ = rng.uniform(0.17, 3.0), rng.uniform(0.17, 3.0)
            shear = rng.uniform(-0.4, 0.4)
            A = np.array([[np.cos(ang) * sx_, np.sin(ang) * sx_],
                          [-np.sin(ang) * sy_ + shear, np.cos(ang) * sy_]])
            # a nearly singular matrix makes the footprint (and the run time of
            # the reference algorithm, CPU or GPU) explode: keep minification
            # within 8x per axis
            if abs(np.linalg.det(A)) > 1e-3 and np.abs(np.linalg.inv(A)).max() <= 8.0:
                break
        persp = rng.uniform(-4e-4, 4e-4, 2) if rng.rand() < 0.4 else (0.0, 0.0)
        M = np.array([[A[0, 0], A[0, 1], persp[0]], [A[1, 0], A[1, 1], persp[1]],
                      [rng.uniform(-10, 30), rng.uniform(-10, 30), 1.0]], np.float32)
        wrap = str(rng.choice(["default", "black", "clamp", "periodic", "mirror"]))
        edgeclamp = bool(rng.rand() < 0.3)
        src_win = None
        if rng.rand() < 0.4:
            src_win = (int(rng.randint(-5, 6)), int(rng.randint(-5, 6)),
                       (0, 0, sw + int(rng.randint(-4, 8)), sh + int(rng.randint(-4, 8))))
        roi = None
        if rng.rand() < 0.3:
            x0, y0 = int(rng.randint(0, dw)), int(rng.randint(0, dh))
            roi = (x0, min(dw, x0 + 1 + int(rng.randint(0, dw))), y0,
                   min(dh, y0 + 1 + int(rng.randint(0, dh))))
        src = synth.image(sw, sh, nch, _np_dtype(st), seed=300 + k)
        pre = synth.image(dw, dh, nch, _np_dtype(dt), seed=600 + k)
        from test_gpu_warp import warp_both
        try:
            g, o = warp_both(oiio, oracle, src, (dh, dw), _np_dtype(dt), M, filt,
                             wrap=wrap, edgeclamp=edgeclamp, src_win=src_win, roi=roi,
                             prefill=pre)
            check(g, o)
        except AssertionError as e:
            raise AssertionError("case %d/%d M=%r filt=%s wrap=%s ec=%s st=%s dt=%s nch=%d "
                                 "src_win=%r roi=%r size=%r->%r: %s"
                                 % (chunk, k, M.tolist(), filt, wrap, edgeclamp, st, dt, nch,
                                    src_win, roi, (sw, sh), (dw, dh), e))


# ----------------------------------------------------- large shapes, banded ----
@pytest.mark.parametrize("chunk", range(3))
def test_resize_fuzz_large(oiio, oracle, chunk):
    """Multi-strip / multi-band plans and the pipelined host path (>= 32 MB):
    the whole image on the GPU, the oracle on a random band of rows."""
    rng = np.random.RandomState(9000 + chunk + SEED)
    for k in range(5):
        sw, sh = int(rng.randint(900, 3300)), int(rng.randint(700, 2600))
        rx = float(rng.choice([0.25, 0.4, 0.5, 0.77, 1.0, 1.6, 2.0]))
        ry = float(rng.choice([0.25, 0.5, 0.5, 0.66, 1.0, 1.5, 2.0]))
        dw, dh = max(8, int(sw * rx)), max(8, int(sh * ry))
        if dw * dh > 12e6:
            dw, dh = dw // 2, dh // 2
        nch = int(rng.choice([1, 3, 4]))
        dt = str(rng.choice(TYPES))
        st = str(rng.choice(TYPES)) if dt == "float" else str(rng.choice([dt, "float"]))
        filt = str(rng.choice(["", "lanczos3", "blackman-harris", "mitchell", "box",
                               "gaussian", "catmull-rom"]))
        src = synth.image(sw, sh, nch, _np_dtype(st), seed=40 + k)
        got = np.zeros((dh, dw, nch), _np_dtype(dt))
        D = oiio.ImageBuf(got)
        assert oiio.ImageBufAlgo.resize(D, oiio.ImageBuf(src), filtername=filt), D.geterror()
        y0 = int(rng.randint(0, max(1, dh - 24)))
        y1 = min(dh, y0 + 24)
        ref = np.zeros_like(got)
        oracle.resize(oracle.Img(ref), oracle.Img(src), filt, roi=(0, dw, y0, y1))
        try:
            check(got[y0:y1], ref[y0:y1])
        except AssertionError as e:
            raise AssertionError("case %d/%d %r: %s" % (chunk, k,
                                 dict(sw=sw, sh=sh, dw=dw, dh=dh, nch=nch, st=st, dt=dt,
                                      filt=filt, band=(y0, y1)), e))


# --------------------------------------------------------------- MIP chain ----
def test_mip_chain_fuzz(oiio, oracle):
    """Odd, prime, 1-pixel-wide and non-square level-0 sizes; box (bit-exact:
    both the 2x2 and the bilinear resize_block paths) and filtered chains."""
    rng = np.random.RandomState(31337 + SEED)
    for k in range(24):
        w = int(rng.choice([1, 2, 3, 5, 16, 33, 64, 100, 127, 255, 256]))
        h = int(rng.choice([1, 2, 4, 7, 32, 50, 64, 99, 128, 200]))
        if w == 1 and h == 1:
            continue
        nch = int(rng.choice([1, 2, 3, 4]))
        filt = str(rng.choice(["box", "box", "lanczos3", "triangle", "blackman-harris"]))
        lvl0 = synth.image(w, h, nch, np.float32, seed=70 + k)
        ref = oracle.mip_chain(lvl0, filt)
        chain = oiio.MipChain(lvl0, filt)
        assert chain.nlevels == len(ref) + 1, (w, h, filt)
        for i, r in enumerate(ref):
            g = chain.download(i + 1)
            assert g.shape == r.shape
            if filt == "box":
                assert np.array_equal(g, r), (w, h, nch, filt, i + 1)
            else:
                assert np.abs(g - r).max() <= 1e-5 * (i + 1), (w, h, nch, filt, i + 1)


# ------------------------------------------------ convolve / fixNonFinite ----
def test_convolve_unsharp_fuzz(oiio, oracle):
    rng = np.random.RandomState(555 + SEED)
    IBA = oiio.ImageBufAlgo
    for k in range(30):
        w, h = int(rng.randint(1, 120)), int(rng.randint(1, 90))
        nch = int(rng.choice([1, 2, 3, 4, 5, 7]))
        st = str(rng.choice(TYPES))
        src = synth.image(w, h, nch, _np_dtype(st), seed=800 + k)
        name = str(rng.choice(["gaussian", "box", "triangle", "mitchell", "binomial", "disk"]))
        kw = int(rng.choice([1, 3, 5, 7, 9]))
        kh = int(rng.choice([1, 3, 5, 7]))
        K, _ = oracle.make_kernel(name, kw, kh)
        swin = None
        if rng.rand() < 0.4:
            swin = (int(rng.randint(-4, 5)), int(rng.randint(-4, 5)),
                    (0, 0, w + int(rng.randint(0, 6)), h + int(rng.randint(0, 6))))
        x, y, full = swin if swin else (0, 0, None)
        ref = np.zeros_like(src)
        got = np.zeros_like(src)
        S, D = oiio.ImageBuf(src), oiio.ImageBuf(got)
        for B in (S, D):
            sp = B.spec()
            sp.x, sp.y = x, y
            (sp.full_x, sp.full_y, sp.full_width, sp.full_height) = \
                full if full else (x, y, w, h)
        if rng.rand() < 0.5:
            oracle.convolve(oracle.Img(ref, x, y, full), oracle.Img(src, x, y, full), K)
            assert IBA.convolve(D, S, IBA.make_kernel(name, kw, kh)), D.geterror()
        else:
            width = float(rng.choice([3.0, 3.0, 5.0]))
            oracle.unsharp_mask(oracle.Img(ref, x, y, full), oracle.Img(src, x, y, full),
                                name if name != "binomial" else "gaussian", width, 0.6, 0.01)
            assert IBA.unsharp_mask(D, S, kernel=name if name != "binomial" else "gaussian",
                                    width=width, contrast=0.6, threshold=0.01), D.geterror()
        assert np.array_equal(got.view(np.uint8), ref.view(np.uint8)), \
            (k, w, h, nch, st, name, kw, kh, swin)


# --------------------------------- device-resident, padded rows, misaligned ----
@pytest.mark.parametrize("chunk", range(4))
def test_resize_fuzz_device_strided(oiio, oracle, chunk):
    """Device tensors that are views into larger allocations: padded row
    strides, first pixel not 16-byte aligned (the vector-load / cp.async paths
    must fall back), canary columns around the destination left untouched."""
    import torch
    rng = np.random.RandomState(6000 + chunk + SEED)
    TT = {"float": torch.float32, "half": torch.float16, "uint8": torch.uint8,
          "uint16": torch.uint16}

    def up(a):
        if a.dtype == np.uint16:
            return torch.from_numpy(a.view(np.int16)).cuda().view(torch.uint16)
        return torch.from_numpy(a).cuda()
    for k in range(30):
        c = _case(rng)
        c["src_win"] = c["dst_win"] = None
        src = synth.image(c["sw"], c["sh"], c["nch"], _np_dtype(c["st"]), seed=k)
        pre = synth.image(c["dw"], c["dh"], c["nch"], _np_dtype(c["dt"]), seed=50 + k)
        ref = pre.copy()
        oracle.resize(oracle.Img(ref), oracle.Img(src), c["filt"], c["fw"], c["roi"])
        padl, padr = int(rng.randint(0, 4)), int(rng.randint(0, 5))
        sbig = torch.zeros((c["sh"], c["sw"] + padl + padr, c["nch"]), dtype=TT[c["st"]],
                           device="cuda")
        sview = sbig[:, padl:padl + c["sw"]]
        sview.copy_(up(src))
        dpl, dpr = int(rng.randint(0, 4)), int(rng.randint(0, 5))
        dbig = torch.zeros((c["dh"], c["dw"] + dpl + dpr, c["nch"]), dtype=TT[c["dt"]],
                           device="cuda")
        if TT[c["dt"]] in (torch.uint8,):
            dbig.fill_(77)
        dview = dbig[:, dpl:dpl + c["dw"]]
        dview.copy_(up(pre))
        before = dbig.clone()
        r = c["roi"]
        roi = oiio.ROI(r[0], r[1], r[2], r[3], 0, 1, 0, c["nch"]) if r else None
        D = oiio.ImageBuf(dview)
        ok = oiio.ImageBufAlgo.resize(D, oiio.ImageBuf(sview), filtername=c["filt"],
                                      filterwidth=c["fw"], roi=roi)
        assert ok, (c, D.geterror())
        got = D.get_pixels()
        try:
            check(np.ascontiguousarray(got), ref, lsb_tol=_int_lsb_tol(oiio, oracle, c, src))
        except AssertionError as e:
            raise AssertionError("case %d/%d %r pads=%r: %s"
                                 % (chunk, k, c, (padl, padr, dpl, dpr), e))
        # canaries: the padding columns are exactly as before
        b8 = lambda t: t.contiguous().view(torch.uint8)
        assert torch.equal(b8(dbig[:, :dpl]), b8(before[:, :dpl]))
        assert torch.equal(b8(dbig[:, dpl + c["dw"]:]), b8(before[:, dpl + c["dw"]:]))


# ----------------------------------------------- batches of frames, one launch ----
@pytest.mark.parametrize("chunk", range(2))
def test_resize_batch_fuzz(oiio, chunk):
    """b2r_resize_batch against one call per frame (which the tests above pin to
    the oracle): random downsizing shapes, types, channel counts and frame counts,
    so the batch plan's band count (searched per batch size) and the per-channel V
    geometry vary; the bytes must not."""
    import torch
    rng = np.random.RandomState(9100 + chunk + SEED)
    TT = {"float": torch.float32, "half": torch.float16, "uint8": torch.uint8,
          "uint16": torch.uint16}
    IBA = oiio.ImageBufAlgo
    for k in range(8):
        st = str(rng.choice(["float", "half", "uint16", "uint8"]))
        dt = st if rng.rand() < 0.7 else "float"
        nch = int(rng.choice([1, 2, 3, 4]))
        sw, sh = int(rng.choice([640, 1000, 1921, 2560])), int(rng.choice([360, 777, 1080]))
        rx, ry = float(rng.choice([0.5, 0.4, 0.66])), float(rng.choice([0.5, 0.4, 0.66]))
        dw, dh = max(2, int(sw * rx)), max(2, int(sh * ry))
        filt = str(rng.choice(["lanczos3", "mitchell", "blackman-harris", "triangle"]))
        nfr = int(rng.choice([2, 3, 7, 19, 64]))
        base = synth.image(sw, sh, nch, _np_dtype(st), seed=40 + k)
        t0 = torch.from_numpy(base.view(np.int16) if st == "uint16" else base).cuda()
        t0 = t0.view(torch.uint16) if st == "uint16" else t0
        srcs = [t0 if i == 0 else
                torch.roll(t0.view(torch.uint8), 5 * i + 1, 0).contiguous().view(t0.dtype)
                for i in range(nfr)]
        one = [torch.zeros((dh, dw, nch), dtype=TT[dt], device="cuda") for _ in range(nfr)]
        many = [torch.zeros_like(t) for t in one]
        for s_, d in zip(srcs, one):
            assert IBA.resize(oiio.ImageBuf(d), oiio.ImageBuf(s_), filtername=filt)
        assert IBA.resize_batch([oiio.ImageBuf(d) for d in many],
                                [oiio.ImageBuf(s_) for s_ in srcs], filtername=filt)
        torch.cuda.synchronize()
        for i, (a, b) in enumerate(zip(one, many)):
            assert torch.equal(a.view(torch.uint8), b.view(torch.uint8)), \
                (chunk, k, i, st, dt, nch, (sw, sh), (dw, dh), filt, nfr)
