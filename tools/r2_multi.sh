#!/bin/bash
# multi-GPU: banded chain with real peer copies + the scaling bench line
N=${1:-2}
O=gpurun_out
timeout 900 python tools/banded_gpu.py 4096 $N 16384 > $O/r2_banded_n$N.log 2>&1; echo "banded rc=$?"
tail -3 $O/r2_banded_n$N.log | cut -c1-2400
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 10 --warmup 3 > $O/r2_scale_n$N.json 2> $O/r2_scale_n$N.err; echo "bench rc=$?"
python - <<PY
import json
d=json.loads(open("$O/r2_scale_n$N.json").read().strip().splitlines()[-1])
print("value", d["value"], "us/frame", d["roofline"]["launch_us"], "e2e", d["e2e"]["value"], "pcie_frac", d["e2e"].get("pcie_frac"), "ceiling/gpu", d["e2e"].get("h2d_ceiling_gbs_per_gpu"))
print("c5", d["configs"]["c5"]["value"], d["configs"]["c5"]["e2e_value"])
print(json.dumps(d.get("banded"), indent=1))
PY
tail -3 $O/r2_scale_n$N.err
