#!/usr/bin/env python
"""Secondary measurements (not the bench.py contract): device-resident MIP
chain (BASELINE config 4) and the other BASELINE configs' kernels.
usage: bench_extra.py mip [size] [filter] | bench_extra.py cfg <c1|c2|c3|c5> [steps]"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def mip(size=16384, filt="box", reps=3):
    import torch
    import openimageio_b200 as oiio
    g = torch.Generator(device="cuda").manual_seed(1)
    lvl0 = torch.rand((size, size, 4), dtype=torch.float32, device="cuda", generator=g)
    best = None
    for _ in range(reps):
        chain = oiio.MipChain(oiio.ImageBuf(lvl0), filt)
        ms = chain.kernel_ms
        n = chain.nlevels
        del chain
        best = ms if best is None else min(best, ms)
    px = sum(max(1, size >> l) ** 2 for l in range(1, n))
    rd = sum(max(1, size >> l) ** 2 for l in range(0, n - 1)) * 16
    wr = px * 16
    print(json.dumps({"workload": "mip %dx%d RGBA f32 %s" % (size, size, filt),
                      "levels": n, "ms": best, "out_Mpix_s": px / best / 1e3,
                      "algorithmic_GBs": (rd + wr) / best / 1e6}))


if __name__ == "__main__":
    if sys.argv[1] == "mip":
        mip(int(sys.argv[2]) if len(sys.argv) > 2 else 16384,
            sys.argv[3] if len(sys.argv) > 3 else "box")
