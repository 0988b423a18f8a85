#!/usr/bin/env python
"""Device-resident 8K RGBA float lanczos3 resize at several downsizing ratios:
which kernel runs and how long it takes."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import openimageio_b200 as oiio
w, h = 7680, 4320
g = torch.Generator(device="cuda").manual_seed(1)
src = torch.rand((h, w, 4), dtype=torch.float32, device="cuda", generator=g)
S = oiio.ImageBuf(src)
host = oiio.ImageBuf(src[:64].cpu().numpy())
for filt in ("lanczos3", "mitchell", "box"):
    for r in (1.5, 2, 3, 4, 6, 8, 16):
        dw, dh = int(w / r), int(h / r)
        dst = torch.empty((dh, dw, 4), dtype=torch.float32, device="cuda")
        D = oiio.ImageBuf(dst)
        for _ in range(2):
            assert oiio.ImageBufAlgo.resize(D, S, filtername=filt)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize(); e0.record()
        for _ in range(5):
            oiio.ImageBufAlgo.resize(D, S, filtername=filt)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 5
        # which path: ask through a small host call of the same geometry ratio
        hd = oiio.ImageBuf(__import__("numpy").zeros((max(1, int(64 / r)), dw, 4), "float32"))
        oiio.ImageBufAlgo.resize(hd, host, filtername=filt)
        rep = oiio.ImageBufAlgo.last_report
        print(json.dumps({"filter": filt, "ratio": r, "ms": round(ms, 4),
                          "GBs": round((src.numel() + dst.numel()) * 4 / ms / 1e6),
                          "path": "fused" if rep.path_used == oiio.PATH_FUSED else "exact",
                          "vmode": rep.fused_vmode, "ht": rep.fused_htaps}))
