import json, os, sys
sys.path.insert(0, "/root/repo")
import torch, numpy as np
import openimageio_b200 as oiio
g = torch.Generator(device="cuda").manual_seed(1)
def run(sw, sh, dw, dh, c, dt, filt):
    tdt = {"f32": torch.float32, "u8": torch.uint8, "f16": torch.float16}[dt]
    if dt == "u8":
        src = torch.randint(0, 256, (sh, sw, c), dtype=tdt, device="cuda", generator=g)
    else:
        src = torch.rand((sh, sw, c), dtype=torch.float32, device="cuda", generator=g).to(tdt)
    dst = torch.empty((dh, dw, c), dtype=tdt, device="cuda")
    S, D = oiio.ImageBuf(src), oiio.ImageBuf(dst)
    for _ in range(2):
        assert oiio.ImageBufAlgo.resize(D, S, filtername=filt), D.geterror()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record()
    for _ in range(5):
        oiio.ImageBufAlgo.resize(D, S, filtername=filt)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    hs = oiio.ImageBuf(src[:max(8, sh // 16)].cpu().numpy())
    hd = oiio.ImageBuf(np.zeros((max(1, dh // 16), dw, c), hs.pixels.dtype))
    hs.spec().full_height = sh; hd.spec().full_height = dh
    oiio.ImageBufAlgo.resize(hd, hs, filtername=filt)
    rep = oiio.ImageBufAlgo.last_report
    gb = (src.numel() * src.element_size() + dst.numel() * dst.element_size()) / ms / 1e6
    print(json.dumps({"case": "%dx%d->%dx%d c%d %s %s" % (sw, sh, dw, dh, c, dt, filt or "default"),
                      "ms": round(ms, 4), "GBs": round(gb),
                      "path": "fused" if rep.path_used == oiio.PATH_FUSED else "exact",
                      "vmode": rep.fused_vmode, "ht": rep.fused_htaps}))
for filt in ("", "lanczos3", "mitchell"):
    for r in (1.25, 2, 3, 4, 8):
        run(1920, 1080, int(1920 * r), int(1080 * r), 4, "f32", filt)
run(3840, 2160, 7680, 1080, 4, "f32", "lanczos3")     # up in x, down in y
run(3840, 2160, 1920, 4320, 4, "f32", "lanczos3")     # down in x, up in y
run(7680, 4320, 1920, 1080, 3, "u8", "lanczos3")
run(7680, 4320, 1920, 1080, 1, "f16", "lanczos3")
run(7680, 4320, 7680, 4320, 4, "f32", "lanczos3")     # 1:1
run(7680, 4320, 7000, 4000, 4, "f32", "")             # slight downsize
