"""Tiled + deflated level output of a size^2 RGBA float chain: where the wall time goes."""
import sys, time, json
import torch
sys.path.insert(0, ".")
import openimageio_b200 as oiio
size = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
g = torch.Generator(device="cuda").manual_seed(7)
lvl0 = torch.rand((size, size, 4), device="cuda", generator=g)
host = torch.empty((size, size, 4), dtype=torch.float32).pin_memory()
host.copy_(lvl0)
torch.cuda.synchronize()
for smooth in (False, True):
    if smooth:  # a picture with structure: deflate finds matches, as it does on real textures
        y = torch.linspace(0, 40, size, device="cuda")[:, None, None]
        x = torch.linspace(0, 40, size, device="cuda")[None, :, None]
        c = torch.arange(4, device="cuda")[None, None, :]
        host.copy_((torch.sin(x + c) * torch.cos(y * 0.7 - c) * 0.5 + 0.5).to(torch.float16).float())
        torch.cuda.synchronize()
    for rep in range(2):
        t0 = time.perf_counter()
        ch = oiio.MipChain(oiio.ImageBuf(host), filtername="lanczos3")
        t1 = time.perf_counter()
        ts = oiio.TileSet(ch, tile=64, half=True, zlevel=1, nthreads=0)
        t2 = time.perf_counter()
        print(json.dumps({"smooth": smooth, "rep": rep, "chain_incl_upload_ms": (t1 - t0) * 1e3,
                          "tileset_wall_ms": (t2 - t1) * 1e3,
                          **{k: ts.stat(k) for k in ("total_ms", "device_ms", "deflate_ms", "setup_ms",
                                                     "wait_ms", "raw_bytes", "stored_bytes")}}))
        del ts, ch
