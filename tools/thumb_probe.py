"""Strong minification (thumbnails): 8K RGBA float down to small sizes, device resident."""
import sys
sys.path.insert(0, ".")
import torch
import openimageio_b200 as oiio
IBA = oiio.ImageBufAlgo
src = torch.rand((4320, 7680, 4), device="cuda")
S = oiio.ImageBuf(src)
for dw, dh in ((480, 270), (320, 180), (256, 144), (128, 72)):
    for filt in ("lanczos3", "box", "triangle"):
        dst = torch.zeros((dh, dw, 4), device="cuda")
        D = oiio.ImageBuf(dst)
        IBA.force_report = True
        IBA.resize(D, S, filtername=filt)
        rep = IBA.last_report
        IBA.force_report = False
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(3):
            IBA.resize(D, S, filtername=filt)
        e1.record(); torch.cuda.synchronize()
        print("8K -> %dx%d %-9s %.3f ms  path %d kernels %d htaps %d" % (dw, dh, filt, e0.elapsed_time(e1) / 3, rep.path_used, rep.kernels_launched, rep.fused_htaps))
