"""The `banded` record of bench.py (N > 1): the 16K^2 RGBA float MIP chain of
BASELINE config 4 cut into row bands over the N GPUs of the box from ONE
process (b2r_mipchain_create_banded), next to the one-GPU chain -- both
INCLUDING the upload of level 0 from pinned host memory, because that is what
banding buys: N PCIe links instead of one.  Rank 0 runs it while the other
ranks wait at a host-side (gloo) barrier, their GPUs idle."""
import time

import numpy as np


def run(H, size, filters=("box", "lanczos3"), devices=None, check=True):
    torch, oiio = H.torch, H.oiio
    n = torch.cuda.device_count() if devices is None else len(devices)
    devices = list(range(n)) if devices is None else list(devices)
    lvl0 = H.c4_level0(size)
    host = torch.empty((size, size, 4), dtype=torch.float32).pin_memory()
    host.copy_(lvl0)
    torch.cuda.synchronize()
    del lvl0
    torch.cuda.empty_cache()
    src = oiio.ImageBuf(host)
    out = {"workload": "maketx MIP chain of a %dx%d RGBA float32 texture, level 0 in "
                       "pinned host memory; wall clock of upload + tables + level loop"
                       % (size, size),
           "n_gpus": len(devices), "unit": "ms"}
    out_px = 0
    w = h = size
    while w > 1 or h > 1:
        w, h = max(1, w // 2), max(1, h // 2)
        out_px += w * h
    for f in filters:
        rec = {}
        # one GPU, upload included (second run: tables cached, pools warm)
        for _ in range(2):
            t0 = time.perf_counter()
            one = oiio.MipChain(src, filtername=f, device=devices[0])
            torch.cuda.synchronize()
            rec["one_gpu_ms"] = (time.perf_counter() - t0) * 1e3
            rec["one_gpu_chain_ms"] = one.kernel_ms
            if _ == 0:
                del one
        for _ in range(2):
            t0 = time.perf_counter()
            many = oiio.MipChainBanded(src, devices, filtername=f)
            rec["banded_ms"] = (time.perf_counter() - t0) * 1e3
            rec["banded_upload_ms"] = many.ms("upload")
            rec["banded_chain_ms"] = many.ms("chain")
            rec["halo_mb"] = many.ms("halo_bytes") / 1e6
            rec["banded_host_ms"] = {"setup": many.ms("setup"), "upload_and_tables": many.ms("phase_a"),
                                     "level_loop": many.ms("phase_b")}
            if _ == 0:
                del many
        rec["speedup"] = rec["one_gpu_ms"] / rec["banded_ms"]
        rec["banded_mpixel_s"] = out_px / 1e6 / (rec["banded_ms"] * 1e-3)
        rec["one_gpu_mpixel_s"] = out_px / 1e6 / (rec["one_gpu_ms"] * 1e-3)
        if check:
            same = True
            for lv in range(2, one.nlevels):
                same = same and bool(np.array_equal(one.download(lv), many.download(lv)))
            # level 1 (a quarter of level 0): compare a few row blocks around
            # the band seams instead of the whole gigabyte
            rec["levels_bit_identical"] = same
        del one, many
        torch.cuda.empty_cache()
        out[f] = rec
    return out
