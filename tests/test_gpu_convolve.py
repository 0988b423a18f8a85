"""GPU parity for make_kernel / convolve / unsharp_mask and maketx's per-level
sharpening: the CUDA kernels through the C ABI against the CPU oracle and the
reference's golden images.  No transcendental runs on the device here, so
everything is held to BIT-EXACTNESS against the oracle."""
import os

import numpy as np
import pytest

import synth
from test_gpu_parity import _np_dtype

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _win(oiio, arr, win):
    B = oiio.ImageBuf(arr)
    if win:
        x, y, full = win
        B.spec().x, B.spec().y = x, y
        (B.spec().full_x, B.spec().full_y, B.spec().full_width,
         B.spec().full_height) = full
    return B


def test_convolve_goldens(oiio, oracle):
    P = np.load(os.path.join(GOLD, "pixelmath_ref.npz"))
    Cv = np.load(os.path.join(GOLD, "conv_ref.npz"))
    IBA = oiio.ImageBufAlgo
    t = oiio.ImageBuf(P["tahoe_small"].copy())
    out = IBA.convolve(t, IBA.make_kernel("bspline", 15, 15))
    assert not out.has_error, out.geterror()
    assert np.array_equal(out.pixels, Cv["bspline_blur"])
    out = IBA.convolve(t, IBA.make_kernel("gaussian", 5, 5))   # oiiotool --blur 5x5
    d = np.abs(out.pixels.astype(int) - Cv["gauss5x5_blur"].astype(int))
    assert d.max() <= 1 and (d != 0).mean() < 2e-5
    out = IBA.unsharp_mask(t)                                   # oiiotool --unsharp
    d = np.abs(out.pixels.astype(int) - Cv["unsharp"].astype(int))
    assert d.max() <= 1 and (d != 0).mean() < 2e-5


@pytest.mark.parametrize("dt,st", [("float", "float"), ("uint8", "uint8"),
                                   ("uint16", "uint16"), ("half", "half"),
                                   ("float", "uint16"), ("uint8", "float")])
@pytest.mark.parametrize("nch", [1, 3, 4, 6])
def test_convolve_bit_exact(oiio, oracle, dt, st, nch):
    src = synth.image(70, 50, nch, _np_dtype(st), seed=70 + nch)
    K, _ = oracle.make_kernel("mitchell", 7, 5)
    ref = np.zeros((50, 70, nch), _np_dtype(dt))
    oracle.convolve(oracle.Img(ref), oracle.Img(src), K)
    got = np.zeros_like(ref)
    Kb = oiio.ImageBufAlgo.make_kernel("mitchell", 7, 5)
    assert oiio.ImageBufAlgo.convolve(oiio.ImageBuf(got), oiio.ImageBuf(src), Kb)
    assert np.array_equal(got.view(np.uint8), ref.view(np.uint8))


def test_convolve_windows_roi_unnormalized(oiio, oracle):
    """A cropped / overscanned source (WrapClamp against the display window,
    black beyond the data window), a dst ROI with a channel sub-range, an
    un-normalised laplacian."""
    src = synth.image(40, 30, 4, np.float32, seed=77)
    swin = (5, -4, (0, 0, 48, 24))
    pre = synth.image(60, 40, 4, np.float32, seed=78)
    for name, w, h, norm in [("gaussian", 5, 5, True), ("laplacian", 3, 3, False),
                             ("binomial", 5, 3, True)]:
        K, _ = oracle.make_kernel(name, w, h)
        ref, got = pre.copy(), pre.copy()
        dwin = (-3, -6, (0, 0, 48, 24))
        oracle.convolve(oracle.Img(ref, *dwin), oracle.Img(src, *swin), K, norm,
                        roi=(0, 50, -2, 30), chrange=(1, 3))
        D = _win(oiio, got, dwin)
        ok = oiio.ImageBufAlgo.convolve(D, _win(oiio, src, swin),
                                        oiio.ImageBufAlgo.make_kernel(name, w, h),
                                        normalize=norm,
                                        roi=oiio.ROI(0, 50, -2, 30, 0, 1, 1, 3))
        assert ok, D.geterror()
        assert np.array_equal(got, ref), name


@pytest.mark.parametrize("st", ["float", "uint8", "half", "uint16"])
@pytest.mark.parametrize("width", [3.0, 7.0])
def test_unsharp_bit_exact(oiio, oracle, st, width):
    """width <= 3: one 2-D pass; width > 3: two 1-D passes (column kernel,
    then its transpose)."""
    src = synth.image(90, 60, 3, _np_dtype(st), seed=81)
    ref = np.zeros_like(src)
    oracle.unsharp_mask(oracle.Img(ref), oracle.Img(src), "gaussian", width, 0.8, 0.02)
    got = np.zeros_like(src)
    D = oiio.ImageBuf(got)
    assert oiio.ImageBufAlgo.unsharp_mask(D, oiio.ImageBuf(src), kernel="gaussian",
                                          width=width, contrast=0.8,
                                          threshold=0.02), D.geterror()
    assert np.array_equal(got.view(np.uint8), ref.view(np.uint8))


def test_unsharp_device_resident_in_place_and_errors(oiio, oracle):
    import torch
    src = synth.image(128, 96, 4, np.float32, seed=83)
    ref = np.zeros_like(src)
    oracle.unsharp_mask(oracle.Img(ref), oracle.Img(src), "gaussian", 3.0, 1.5, 0.0)
    t = torch.from_numpy(src).cuda()
    B = oiio.ImageBuf(t)
    assert oiio.ImageBufAlgo.unsharp_mask(B, B, contrast=1.5)      # in place
    assert np.array_equal(t.cpu().numpy(), ref)
    IBA = oiio.ImageBufAlgo
    bad = IBA.unsharp_mask(oiio.ImageBuf(src), kernel="median")
    assert bad.has_error
    bad = IBA.unsharp_mask(oiio.ImageBuf(src), kernel="nope")
    assert bad.geterror() == 'Unknown kernel "nope" 3x3'
    D = oiio.ImageBuf(np.zeros((96, 128, 3), np.float32))
    assert not IBA.convolve(D, oiio.ImageBuf(src), IBA.make_kernel("box", 3, 3))
    assert D.geterror() == "images must have the same number of channels"


@pytest.mark.parametrize("first", [False, True])
@pytest.mark.parametrize("filt", ["lanczos3", "box"])
def test_mip_chain_sharpen(oiio, oracle, filt, first):
    """maketx --sharpen 0.6 [--attrib maketx:sharpenfirst 1]; with sharpening
    the box fast path is off (maketexture.cpp:880-881)."""
    lvl0 = synth.image(96, 64, 4, np.float32, seed=121)
    ref = oracle.mip_chain(lvl0, filt, sharpen=0.6, sharpen_first=first)
    chain = oiio.MipChain(lvl0, filt, sharpen=0.6, sharpen_first=first)
    assert chain.nlevels == len(ref) + 1
    assert np.array_equal(chain.download(0), lvl0)
    for i, r in enumerate(ref):
        g = chain.download(i + 1)
        assert np.abs(g - r).max() <= 2e-5 * (i + 1), "level %d" % (i + 1)
    plain = oiio.MipChain(lvl0, filt).download(1)
    assert not np.array_equal(plain, chain.download(1))


def test_mip_chain_sharpen_with_highlightcomp(oiio, oracle):
    lvl0 = (synth.image(96, 64, 4, np.float32, seed=122) ** 4 * np.float32(30)).astype(np.float32)
    for first in (False, True):
        ref = oracle.mip_chain(lvl0, "lanczos3", highlightcomp=True, alpha_channel=3,
                               sharpen=0.5, sharpen_first=first)
        chain = oiio.MipChain(lvl0, "lanczos3", highlightcomp=True, sharpen=0.5,
                              sharpen_first=first)
        for i, r in enumerate(ref):
            g = chain.download(i + 1)
            assert g.min() >= 0.0                       # clamp after rangeexpand
            scale = max(1.0, float(np.abs(r).max()))
            assert np.abs(g - r).max() <= 1e-4 * (i + 1) * scale, "level %d" % (i + 1)
