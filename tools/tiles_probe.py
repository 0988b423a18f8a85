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
    for rep, pred in ((0, 0), (1, 0), (2, 3), (3, 3)):
        t0 = time.perf_counter()
        ch = oiio.MipChain(oiio.ImageBuf(host), filtername="lanczos3")
        t1 = time.perf_counter()
        ts = oiio.TileSet(ch, tile=64, half=True, zlevel=1, nthreads=0, predictor=pred)
        t2 = time.perf_counter()
        rec = {"smooth": smooth, "rep": rep, "predictor": pred,
               "chain_incl_upload_ms": (t1 - t0) * 1e3, "tileset_wall_ms": (t2 - t1) * 1e3,
               **{k: ts.stat(k) for k in ("total_ms", "device_ms", "deflate_ms", "setup_ms",
                                          "wait_ms", "raw_bytes", "stored_bytes")}}
        if rep == 3:  # the texture file itself (tmpfs / local disk of the box)
            t3 = time.perf_counter()
            ts.write_tiff("/tmp/tiles_probe.tif")
            rec["write_tiff_ms"] = (time.perf_counter() - t3) * 1e3
            import os
            rec["file_mb"] = os.path.getsize("/tmp/tiles_probe.tif") / 1e6
            os.remove("/tmp/tiles_probe.tif")
        print(json.dumps(rec))
        del ts, ch
# the re-layout kernels alone (device in, device out): CUDA events around 20 launches
for half in (False, True):
    for pred in (0, 3):
        for _ in range(3):
            out = oiio.tile_relayout(lvl0, tile=64, half=half, predictor=pred)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record()
        for _ in range(20):
            out = oiio.tile_relayout(lvl0, tile=64, half=half, predictor=pred)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 20
        by = size * size * 4 * (4 + (2 if half else 4))
        print(json.dumps({"relayout_kernel": "half" if half else "float", "predictor": pred,
                          "ms": ms, "algorithmic_gbs": by / ms / 1e6}))
