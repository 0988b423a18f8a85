import sys, json
sys.path.insert(0,"/root/repo"); sys.path.insert(0,"/root/repo/tests")
import torch, numpy as np, openimageio_b200 as oiio
w,h,c=7680,4320,6
g=torch.Generator(device="cuda").manual_seed(1)
src=torch.rand((h,w,c),dtype=torch.float32,device="cuda",generator=g)
dst=torch.empty((h//2,w//2,c),dtype=torch.float32,device="cuda")
S,D=oiio.ImageBuf(src),oiio.ImageBuf(dst)
def t(n=5,**kw):
    for _ in range(2): assert oiio.ImageBufAlgo.resize(D,S,filtername="lanczos3",**kw)
    e0,e1=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record()
    for _ in range(n): oiio.ImageBufAlgo.resize(D,S,filtername="lanczos3",**kw)
    e1.record(); torch.cuda.synchronize(); return e0.elapsed_time(e1)/n
split=t(); a=dst.clone()
exact=t(n=2,path=oiio.PATH_EXACT)
print(json.dumps({"six_channel_8k_split_ms":split,"exact_ms":exact,"max_abs_diff":float((a-dst).abs().max())}))
