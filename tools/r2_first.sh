#!/bin/bash
# round 2, first GPU contact of the stream kernel: smoke, parity, A/B bench
O=gpurun_out
mkdir -p $O
timeout 180 python -c "
import __graft_entry__ as g
g.smoke()
" > $O/r2_smoke.log 2>&1; echo "smoke rc=$?" | tee -a $O/r2_smoke.log
tail -3 $O/r2_smoke.log
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu > $O/r2_parity.log 2>&1; echo "parity rc=$?"
tail -5 $O/r2_parity.log
for wl in c0 c1 c2 c5; do
  for s in 1 0; do
    B2R_STREAM=$s timeout 300 python bench.py --workload $wl --steps 25 --no-cpu --e2e-steps 3 > $O/r2_ab_${wl}_s${s}.json 2> $O/r2_ab_${wl}_s${s}.err
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
