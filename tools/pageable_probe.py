#!/usr/bin/env python
"""End-to-end resize from PAGEABLE host memory (what an ImageBuf normally
owns) against pinned host memory, C0 shape."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import openimageio_b200 as oiio
import synth

sw, sh, dw, dh = 7680, 4320, 3840, 2160
src = synth.image(sw, sh, 4, np.float32)
dst = np.zeros((dh, dw, 4), np.float32)
IBA = oiio.ImageBufAlgo


def run(S, D, n=5):
    for _ in range(2):
        assert IBA.resize(D, S, filtername="lanczos3")
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(n):
        assert IBA.resize(D, S, filtername="lanczos3")
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / n * 1e3


ms_page = run(oiio.ImageBuf(src), oiio.ImageBuf(dst))
ps = torch.from_numpy(src).pin_memory()
pd = torch.empty((dh, dw, 4), dtype=torch.float32).pin_memory()
ms_pin = run(oiio.ImageBuf(ps), oiio.ImageBuf(pd))
same = bool(np.array_equal(dst, pd.numpy()))
print(json.dumps({"pageable_ms": ms_page, "pinned_ms": ms_pin,
                  "pageable_Mpix_s": dw * dh / ms_page / 1e3,
                  "pinned_Mpix_s": dw * dh / ms_pin / 1e3, "same_result": same,
                  "cores": os.cpu_count()}))

# other entry points with pageable host images: upload / download rates
# (second call of each: the pinned rings and the device pool are warm)
big = synth.image(8192, 4096, 4, np.float32, seed=3)          # 537 MB
for rep in range(2):
    t0 = time.perf_counter()
    chain = oiio.MipChain(big, "box")
    torch.cuda.synchronize()
    t_up = time.perf_counter() - t0
    t0 = time.perf_counter()
    l1 = chain.download(1)
    t_dn = time.perf_counter() - t0
    if rep == 0:
        del chain
ref = oiio.MipChain(torch.from_numpy(big).cuda(), "box").download(1)
out = np.zeros_like(big)
for rep in range(2):
    t0 = time.perf_counter()
    assert IBA.unsharp_mask(oiio.ImageBuf(out), oiio.ImageBuf(big))
    t_us = time.perf_counter() - t0
pin = torch.from_numpy(big).pin_memory()
pout = torch.empty_like(pin).pin_memory()
for rep in range(2):
    t0 = time.perf_counter()
    assert IBA.unsharp_mask(oiio.ImageBuf(pout), oiio.ImageBuf(pin))
    t_usp = time.perf_counter() - t0
print(json.dumps({"mipchain_create_pageable_537MB_ms": t_up * 1e3,
                  "download_level1_134MB_ms": t_dn * 1e3,
                  "download_GBs": l1.nbytes / t_dn / 1e9,
                  "level1_same_as_device_chain": bool(np.array_equal(l1, ref)),
                  "unsharp_pageable_in_out_537MB_ms": t_us * 1e3,
                  "unsharp_pinned_in_out_537MB_ms": t_usp * 1e3,
                  "unsharp_same": bool(np.array_equal(out, pout.numpy()))}))
