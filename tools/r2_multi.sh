#!/bin/bash
# multi-GPU: banded chain with real peer copies + the scaling bench line
N=${1:-2}
O=gpurun_out
timeout 900 python tools/banded_gpu.py 4096 $N 16384 > $O/r2_banded_n$N.log 2>&1; echo "banded rc=$?"
tail -3 $O/r2_banded_n$N.log | cut -c1-1800
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 10 --warmup 3 > $O/r2_scale_n$N.json 2> $O/r2_scale_n$N.err; echo "bench rc=$?"
tail -c 2500 $O/r2_scale_n$N.json; tail -3 $O/r2_scale_n$N.err
