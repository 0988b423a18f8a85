"""Multi-GPU check + timing of the banded MIP chain (run under gpurun --gpus N):
every level bit-identical to the single-device chain, with real peer copies.
usage: python tools/banded_gpu.py [size] [ngpus]"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch  # noqa: E402

import openimageio_b200 as oiio  # noqa: E402
import synth  # noqa: E402

size = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
n = int(sys.argv[2]) if len(sys.argv) > 2 else torch.cuda.device_count()
res = {"size": size, "n": n}
for filt in ("box", "lanczos3"):
    src = synth.image(size, size - 37, 4, np.float32, seed=77)
    one = oiio.MipChain(src, filtername=filt)
    many = oiio.MipChainBanded(src, list(range(n)), filtername=filt)
    ok = all(np.array_equal(one.download(lv), many.download(lv)) for lv in range(one.nlevels))
    res[filt] = {"bit_identical": bool(ok), "halo_mb": many.ms("halo_bytes") / 1e6,
                 "chain_ms": many.ms("chain"), "one_gpu_chain_ms": one.kernel_ms}
    del one, many
print(json.dumps(res))


class H:  # the bits of bench.Harness banded_bench needs
    pass


import bench  # noqa: E402
import banded_bench  # noqa: E402
h = bench.Harness.__new__(bench.Harness)
h.torch, h.oiio, h.dev = torch, oiio, torch.device("cuda", 0)
h.local_rank, h.rank, h.world = 0, 0, 1
big = int(sys.argv[3]) if len(sys.argv) > 3 else 16384
print(json.dumps(banded_bench.run(h, big, devices=list(range(n)))))
