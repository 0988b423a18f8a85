import sys, json
sys.path.insert(0, ".")
import torch
import openimageio_b200 as oiio
IBA = oiio.ImageBufAlgo
def timed(fn, n=10):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
for nch in (1, 2, 3, 4):
    for dt in (torch.float32, torch.float16, torch.uint8):
        if dt == torch.uint8:
            src = torch.randint(0, 256, (4320, 7680, nch), dtype=dt, device="cuda")
        else:
            src = torch.rand((4320, 7680, nch), device="cuda").to(dt)
        dst = torch.zeros((2160, 3840, nch), dtype=dt, device="cuda")
        S, D = oiio.ImageBuf(src), oiio.ImageBuf(dst)
        ms = timed(lambda: IBA.resize(D, S, filtername="lanczos3"))
        nbytes = src.numel() * src.element_size() + dst.numel() * dst.element_size()
        print("8K->4K lanczos3 nch=%d %-14s %.3f ms  %.0f GB/s" % (nch, str(dt), ms, nbytes / ms / 1e6))
