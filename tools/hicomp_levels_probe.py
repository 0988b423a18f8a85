"""Per-level cost of highlight compensation in a 16K^2 chain: plain vs hicomp resize of
each level's shape, device resident (which levels take the fused kernel)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import openimageio_b200 as oiio
IBA = oiio.ImageBufAlgo


def timed(fn, n=10):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


size = 16384
while size >= 512:
    src = torch.rand((size, size, 4), device="cuda") * 4
    dst = torch.zeros((size // 2, size // 2, 4), device="cuda")
    S, D = oiio.ImageBuf(src), oiio.ImageBuf(dst)
    S.spec().alpha_channel = D.spec().alpha_channel = 3
    p = timed(lambda: IBA.resize(D, S, filtername="lanczos3"))
    h = timed(lambda: IBA.resize(D, S, filtername="lanczos3", highlightcomp=True))
    print("%5d -> %5d: plain %.3f ms, hicomp %.3f ms" % (size, size // 2, p, h))
    del src, dst, S, D
    size //= 2
