import torch, time
n=530841600
h=torch.empty(n,dtype=torch.uint8).pin_memory()
d=torch.empty(n,dtype=torch.uint8,device='cuda')
for _ in range(3): d.copy_(h,non_blocking=True)
torch.cuda.synchronize()
e0=torch.cuda.Event(enable_timing=True);e1=torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10): d.copy_(h,non_blocking=True)
e1.record();torch.cuda.synchronize()
ms=e0.elapsed_time(e1)/10
print("H2D 531MB: %.3f ms  %.1f GB/s"%(ms,n/ms/1e6))
# split in 16 chunks on a second stream
s=torch.cuda.Stream()
h2=torch.empty(132710400,dtype=torch.uint8).pin_memory()
d2=torch.empty(132710400,dtype=torch.uint8,device='cuda')
e0.record()
for _ in range(10):
    d.copy_(h,non_blocking=True)
    with torch.cuda.stream(s):
        h2.copy_(d2,non_blocking=True)
e1.record();s.synchronize();torch.cuda.synchronize()
ms=e0.elapsed_time(e1)/10
print("H2D 531MB with concurrent D2H 133MB: %.3f ms  %.1f GB/s"%(ms,n/ms/1e6))
