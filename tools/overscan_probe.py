import sys, json
sys.path.insert(0,"/root/repo")
import torch, openimageio_b200 as oiio
w,h=7680,4320
g=torch.Generator(device="cuda").manual_seed(1)
src=torch.rand((h,w,4),dtype=torch.float32,device="cuda",generator=g)
dst=torch.empty((h//2,w//2,4),dtype=torch.float32,device="cuda")
S,D=oiio.ImageBuf(src),oiio.ImageBuf(dst)
sp=S.spec(); sp.x,sp.y=-40,-30; sp.full_x=sp.full_y=0; sp.full_width,sp.full_height=w-80,h-60
dp=D.spec(); dp.x,dp.y=-20,-15; dp.full_x=dp.full_y=0; dp.full_width,dp.full_height=(w-80)//2,(h-60)//2
def t(n=5,**kw):
    for _ in range(2): assert oiio.ImageBufAlgo.resize(D,S,filtername="lanczos3",**kw), D.geterror()
    e0,e1=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record()
    for _ in range(n): oiio.ImageBufAlgo.resize(D,S,filtername="lanczos3",**kw)
    e1.record(); torch.cuda.synchronize(); return e0.elapsed_time(e1)/n
print(json.dumps({"overscan_8k_split_ms":t(),"overscan_8k_exact_ms":t(n=2,path=oiio.PATH_EXACT)}))
