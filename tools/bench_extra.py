#!/usr/bin/env python
"""Secondary measurements (not the bench.py contract): device-resident MIP
chain (BASELINE config 4) and the other BASELINE configs' kernels.
usage: bench_extra.py mip [size] [filter] [hicomp]
       bench_extra.py warp         rotate / resize:from-to / fit exact, 4K RGBA f32 and u8
       bench_extra.py conv         convolve / unsharp_mask / MIP chain with --sharpen
       bench_extra.py maketx       pixel statistics, fixNonFinite, fillholes_pushpull
       bench_extra.py range        rangecompress / rangeexpand / highlightcomp resize, 8K RGBA f32"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def mip(size=16384, filt="box", reps=3, hicomp=False):
    import torch
    import openimageio_b200 as oiio
    g = torch.Generator(device="cuda").manual_seed(1)
    lvl0 = torch.rand((size, size, 4), dtype=torch.float32, device="cuda", generator=g)
    best = None
    for _ in range(reps):
        chain = oiio.MipChain(oiio.ImageBuf(lvl0), filt, highlightcomp=hicomp)
        ms = chain.kernel_ms
        n = chain.nlevels
        del chain
        best = ms if best is None else min(best, ms)
    px = sum(max(1, size >> l) ** 2 for l in range(1, n))
    rd = sum(max(1, size >> l) ** 2 for l in range(0, n - 1)) * 16
    wr = px * 16
    print(json.dumps({"workload": "mip %dx%d RGBA f32 %s%s" % (size, size, filt,
                                                                " hicomp" if hicomp else ""),
                      "levels": n, "ms": best, "out_Mpix_s": px / best / 1e3,
                      "algorithmic_GBs": (rd + wr) / best / 1e6}))


def _time(fn, steps=20, warmup=3):
    import torch
    for _ in range(warmup):
        fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(steps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps


def range_ops():
    """Streaming point ops: algorithmic bytes = read src + write dst."""
    import torch
    import openimageio_b200 as oiio
    w, h = 7680, 4320
    g = torch.Generator(device="cuda").manual_seed(2)
    src = torch.rand((h, w, 4), dtype=torch.float32, device="cuda", generator=g) * 8
    dst = torch.empty_like(src)
    S, D = oiio.ImageBuf(src), oiio.ImageBuf(dst)
    for name in ("rangecompress", "rangeexpand"):
        fn = getattr(oiio.ImageBufAlgo, name)
        for luma in (False, True):
            ms = _time(lambda: fn(D, S, useluma=luma))
            print(json.dumps({"workload": "%s 8K RGBA f32 luma=%d" % (name, luma),
                              "ms": ms, "algorithmic_GBs": 2 * src.numel() * 4 / ms / 1e6}))
    small = torch.empty((h // 2, w // 2, 4), dtype=torch.float32, device="cuda")
    R = oiio.ImageBuf(small)
    for hc in (False, True):
        ms = _time(lambda: oiio.ImageBufAlgo.resize(R, S, filtername="lanczos3",
                                                    highlightcomp=hc))
        print(json.dumps({"workload": "8K->4K lanczos3 RGBA f32 highlightcomp=%d" % hc,
                          "ms": ms}))


def warp_ops():
    """warp kernel, device resident.  Reported per output pixel; the kernel
    is FP32/L1 bound (taps x channels FMAs per pixel), not HBM bound."""
    import math
    import torch
    import openimageio_b200 as oiio
    IBA = oiio.ImageBufAlgo
    w, h = 3840, 2160
    g = torch.Generator(device="cuda").manual_seed(3)
    for dt in (torch.float32, torch.float16, torch.uint8):
        if dt == torch.uint8:
            src = torch.randint(0, 256, (h, w, 4), dtype=dt, device="cuda", generator=g)
        else:
            src = torch.rand((h, w, 4), dtype=torch.float32, device="cuda",
                             generator=g).to(dt)
        dst = torch.empty_like(src)
        S, D = oiio.ImageBuf(src), oiio.ImageBuf(dst)
        M = oiio.rotate_matrix(30.0 * math.pi / 180.0, w / 2, h / 2)
        for filt in ("lanczos3", "mitchell", "triangle"):
            ms = _time(lambda: IBA.warp(D, S, M, filtername=filt, wrap="black"))
            IBA.contract_fma = True   # tolerance mode: packed FMAs in the accumulation
            ms_fma = _time(lambda: IBA.warp(D, S, M, filtername=filt, wrap="black"))
            IBA.contract_fma = False
            print(json.dumps({"workload": "rotate 30deg 4K RGBA %s %s" % (dt, filt),
                              "ms": ms, "out_Mpix_s": w * h / ms / 1e3,
                              "ms_contract_fma": ms_fma}))
        # 2x minification through the warp path (oiiotool --resize:from=...)
        small = torch.empty((h // 2, w // 2, 4), dtype=dt, device="cuda")
        R = oiio.ImageBuf(small)
        Mh = oiio.fromto_matrix((0.5, 0.5, w - 1, h - 1), (0, 0, w // 2, h // 2))
        ms = _time(lambda: IBA.warp(R, S, Mh, filtername="lanczos3", edgeclamp=True))
        print(json.dumps({"workload": "warp 4K->2K (from/to) RGBA %s lanczos3" % dt,
                          "ms": ms, "out_Mpix_s": w * h / 4 / ms / 1e3}))


def conv_ops():
    """convolve / unsharp_mask, 8K RGBA f32 device resident.  Algorithmic
    bytes: convolve = src + dst; unsharp = src + dst (the blur image is a
    device temporary: + one write and one read of it in this version)."""
    import torch
    import openimageio_b200 as oiio
    IBA = oiio.ImageBufAlgo
    w, h = 7680, 4320
    g = torch.Generator(device="cuda").manual_seed(4)
    src = torch.rand((h, w, 4), dtype=torch.float32, device="cuda", generator=g)
    dst = torch.empty_like(src)
    S, D = oiio.ImageBuf(src), oiio.ImageBuf(dst)
    nbytes = 2 * src.numel() * 4
    for name, kw in (("gaussian", 3), ("gaussian", 5), ("bspline", 15)):
        K = IBA.make_kernel(name, kw, kw)
        ms = _time(lambda: IBA.convolve(D, S, K))
        print(json.dumps({"workload": "convolve %s %dx%d 8K RGBA f32" % (name, kw, kw),
                          "ms": ms, "algorithmic_GBs": nbytes / ms / 1e6,
                          "GFMA_s": w * h * 4 * kw * kw / ms / 1e6}))
    for width in (3.0, 7.0):
        ms = _time(lambda: IBA.unsharp_mask(D, S, width=width, contrast=0.6))
        print(json.dumps({"workload": "unsharp_mask gaussian width %g 8K RGBA f32" % width,
                          "ms": ms, "algorithmic_GBs": nbytes / ms / 1e6}))
    lvl0 = torch.rand((8192, 8192, 4), dtype=torch.float32, device="cuda", generator=g)
    for sh in (0.0, 0.6):
        best = None
        for _ in range(3):
            chain = oiio.MipChain(oiio.ImageBuf(lvl0), "lanczos3", sharpen=sh)
            best = chain.kernel_ms if best is None else min(best, chain.kernel_ms)
            del chain
        print(json.dumps({"workload": "mip 8192^2 RGBA f32 lanczos3 sharpen=%g" % sh,
                          "ms": best}))


def maketx_ops():
    """The whole-image maketx stages either side of the MIP chain, 8K RGBA f32
    device resident: one read of the image each (stats), read + write (fixnan)."""
    import torch
    import openimageio_b200 as oiio
    IBA = oiio.ImageBufAlgo
    w, h = 7680, 4320
    g = torch.Generator(device="cuda").manual_seed(5)
    src = torch.rand((h, w, 4), dtype=torch.float32, device="cuda", generator=g)
    src[100, 100, 2] = float("nan")
    S = oiio.ImageBuf(src)
    nbytes = src.numel() * 4
    ms = _time(lambda: IBA.computePixelStats(S))
    print(json.dumps({"workload": "computePixelStats+isMonochrome 8K RGBA f32 (incl. host merge)",
                      "ms": ms, "algorithmic_GBs": nbytes / ms / 1e6}))
    dst = torch.empty_like(src)
    D = oiio.ImageBuf(dst)
    for mode, name in ((oiio.NONFINITE_BOX3, "box3"), (oiio.NONFINITE_NONE, "count only")):
        ms = _time(lambda: IBA.fixNonFinite(D, S, mode))
        print(json.dumps({"workload": "fixNonFinite %s 8K RGBA f32" % name, "ms": ms,
                          "algorithmic_GBs": 2 * nbytes / ms / 1e6}))
    hole = torch.rand((2160, 3840, 4), dtype=torch.float32, device="cuda", generator=g)
    hole[..., 3] = (hole[..., 3] > 0.3).float()
    hole[..., :3] *= hole[..., 3:4]
    H, O = oiio.ImageBuf(hole), oiio.ImageBuf(torch.empty_like(hole))
    ms = _time(lambda: IBA.fillholes_pushpull(O, H), steps=5)
    print(json.dumps({"workload": "fillholes_pushpull 4K RGBA f32 (12 levels down + up)",
                      "ms": ms}))


if __name__ == "__main__":
    if sys.argv[1] == "mip":
        mip(int(sys.argv[2]) if len(sys.argv) > 2 else 16384,
            sys.argv[3] if len(sys.argv) > 3 else "box",
            hicomp=len(sys.argv) > 4 and sys.argv[4] == "hicomp")
    elif sys.argv[1] == "range":
        range_ops()
    elif sys.argv[1] == "warp":
        warp_ops()
    elif sys.argv[1] == "conv":
        conv_ops()
    elif sys.argv[1] == "maketx":
        maketx_ops()
