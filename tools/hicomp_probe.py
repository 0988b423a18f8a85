"""8K -> 4K lanczos3 RGBA float, device resident: plain resize, resize with
highlight compensation (fused into the stream kernel), and the three separate
passes (rangecompress, resize, rangeexpand) it replaces."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np, torch
import openimageio_b200 as oiio, synth
IBA = oiio.ImageBufAlgo
src = torch.from_numpy((synth.image(7680, 4320, 4, np.float32, seed=5) ** 3 * 12).astype(np.float32)).cuda()
dst = torch.zeros((2160, 3840, 4), dtype=torch.float32, device="cuda")
tmp = torch.zeros_like(src)
S, D, T = oiio.ImageBuf(src), oiio.ImageBuf(dst), oiio.ImageBuf(tmp)
for b in (S, D, T):
    b.spec().alpha_channel = 3


def timed(fn, n=20):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


def three():
    IBA.rangecompress(T, S)
    IBA.resize(D, T, filtername="lanczos3")
    IBA.rangeexpand(D, D)


print("plain  resize        %.3f ms" % timed(lambda: IBA.resize(D, S, filtername="lanczos3")))
print("highlightcomp (one call) %.3f ms" % timed(lambda: IBA.resize(D, S, filtername="lanczos3", highlightcomp=True)))
a = dst.clone()
print("three separate passes %.3f ms" % timed(three))
print("max |fused - three passes| = %.3g" % float((a - dst).abs().max()))
# error of the fused call (lg2 / ex2 on the special-function unit) against the three
# passes (libm logf / expf): relative to the value range, and per value where the
# curves act (|v| > 0.18; below that they are the identity)
d = (a - dst).abs()
print("max |fused - three passes| / max|v| = %.3g (values up to %.3g)"
      % (float(d.max() / dst.abs().max()), float(dst.abs().max())))
m = dst.abs() > 0.18
print("max relative difference over |v| > 0.18 = %.3g" % float((d[m] / dst.abs()[m]).max()))
big = torch.rand((16384, 16384, 4), device="cuda")
for hc in (False, True):
    ch = oiio.MipChain(oiio.ImageBuf(big), "lanczos3", highlightcomp=hc)
    ch = oiio.MipChain(oiio.ImageBuf(big), "lanczos3", highlightcomp=hc)
    print("16K^2 lanczos3 chain, hicomp=%s: %.3f ms" % (hc, ch.kernel_ms))
    del ch
