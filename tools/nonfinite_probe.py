import sys, os, json
sys.path.insert(0,"/root/repo"); sys.path.insert(0,"/root/repo/tests")
import torch, numpy as np, openimageio_b200 as oiio
w,h=7680,4320
g=torch.Generator(device="cuda").manual_seed(1)
src=torch.rand((h,w,4),dtype=torch.float32,device="cuda",generator=g)
dst=torch.empty((h//2,w//2,4),dtype=torch.float32,device="cuda")
S,D=oiio.ImageBuf(src),oiio.ImageBuf(dst)
def t(n=10):
    for _ in range(3): oiio.ImageBufAlgo.resize(D,S,filtername="lanczos3")
    e0,e1=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record()
    for _ in range(n): oiio.ImageBufAlgo.resize(D,S,filtername="lanczos3")
    e1.record(); torch.cuda.synchronize(); return e0.elapsed_time(e1)/n
clean=t()
src[1000,2000,1]=float("inf")
one=t()
src[::97,::89,2]=float("nan")
many=t()
# dense: 1 % of all values NaN / +-Inf (nearly every output pixel is non-finite in the reference too)
m=torch.rand((h,w,4),device="cuda",generator=g)<0.01
src[m]=float("nan")
dense=t()
print(json.dumps({"clean_ms":clean,"one_inf_ms":one,"scattered_nan_ms":many,"dense_1pct_nan_ms":dense}))
