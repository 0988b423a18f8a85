#!/bin/bash
O=gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_maketx.py -x -q -m gpu > $O/r2_parity.log 2>&1; echo "parity rc=$?"
tail -3 $O/r2_parity.log
for wl in c0 c1 c5; do
  for s in 1; do
    B2R_STREAM=$s timeout 300 python bench.py --workload $wl --steps 25 --no-cpu --e2e-steps 3 --no-configs --sustain-ms 0 > $O/r2_ab_${wl}_s${s}.json 2> $O/r2_ab_${wl}_s${s}.err
    python - <<PY
import json
try:
    d=json.loads(open("$O/r2_ab_${wl}_s${s}.json").read().strip().splitlines()[-1])
    print("$wl stream=$s", "us/frame %.1f" % d["roofline"]["launch_us"], "frac %.3f" % d["roofline"]["frac"], "same", d["device_equals_host_result"], d["clocks"]["sm_mhz"], d["clocks"]["reasons"])
except Exception as e:
    print("$wl stream=$s failed", e)
PY
  done
done
bash tools/r2_prof.sh f c0
