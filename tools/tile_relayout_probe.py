"""The tile re-layout kernels on their own (b2r_tile_relayout: device in, device out):
CUDA-event time per launch and algorithmic GB/s for a size^2 RGBA float level.
usage: tile_relayout_probe.py [size] [launches]"""
import json
import sys

import torch

sys.path.insert(0, ".")
import openimageio_b200 as oiio

size = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
n = int(sys.argv[2]) if len(sys.argv) > 2 else 20
g = torch.Generator(device="cuda").manual_seed(7)
lvl0 = torch.rand((size, size, 4), device="cuda", generator=g)
for half in (False, True):
    for pred in (0, 3):
        for _ in range(3):
            out = oiio.tile_relayout(lvl0, tile=64, half=half, predictor=pred)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record()
        for _ in range(n):
            out = oiio.tile_relayout(lvl0, tile=64, half=half, predictor=pred)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / n
        by = size * size * 4 * (4 + (2 if half else 4))
        print(json.dumps({"relayout_kernel": "half" if half else "float", "predictor": pred,
                          "ms": ms, "algorithmic_gbs": by / ms / 1e6}))
