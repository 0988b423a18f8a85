#!/usr/bin/env python
"""Row-band MIP chain over N GPUs (one process per GPU):
   python -m torch.distributed.run --nproc-per-node N tools/mipdist_gpu.py [size] [filter]
Checks every banded level bit for bit against the single-GPU chain on rank 0
and prints the device time of the banded part."""
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import openimageio_b200 as oiio  # noqa: E402
from openimageio_b200 import mipdist  # noqa: E402


def main():
    size = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
    filt = sys.argv[2] if len(sys.argv) > 2 else "lanczos3"
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    rank, world = dist.get_rank(), dist.get_world_size()
    g = torch.Generator(device="cpu").manual_seed(7)
    check = size <= 8192
    y0, y1 = mipdist.split_rows(size, world, rank)
    if check:
        lvl0 = torch.rand((size, size, 4), generator=g)          # same on all ranks
        band = lvl0[y0:y1].to(dev)
    else:
        band = torch.rand((y1 - y0, size, 4), device=dev)
    # warm-up (plans, NCCL channels), then a timed run
    for it in range(2):
        torch.cuda.synchronize()
        dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        chain = mipdist.DistributedMipChain(band, size, size, filt)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
    t = torch.tensor([ms], device=dev, dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ok = True
    nb = len(chain.bands)
    if check:
        ref = None
        if rank == 0:
            one = oiio.MipChain(oiio.ImageBuf(lvl0.to(dev)), filt)
            ref = [one.download(i) for i in range(one.nlevels)]
        for k in range(1, nb):
            full = chain.gather_level(k)
            if rank == 0:
                ok = ok and np.array_equal(full.cpu().numpy(), ref[k])
        if rank == 0:
            for i, tt in enumerate(chain.tail):
                a, b = tt.cpu().numpy(), ref[nb + i]
                ok = ok and (np.array_equal(a, b) if filt == "box"
                             else np.abs(a - b).max() <= 1e-5 * (nb + i))
    if rank == 0:
        print(json.dumps({"workload": "banded mip %dx%d RGBA f32 %s" % (size, size, filt),
                          "n_gpus": world, "levels": chain.nlevels,
                          "banded_levels": nb - 1, "ms_max_over_ranks": float(t.item()),
                          "halo_bytes_rank0": chain.halo_bytes,
                          "matches_single_gpu": bool(ok) if check else None}))
    dist.destroy_process_group()
    return 0 if ok else 1


if __name__ == "__main__":
    sys.exit(main())
