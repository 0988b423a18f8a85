#!/bin/bash
# whole GPU suite + per-config device numbers + (FULL=1) the default bench line
O=gpurun_out
T=${TAG:-x}
timeout 2400 python -m pytest tests -x -q -m gpu > $O/r2_gpu_tests.log 2>&1; echo "gpu tests rc=$?"
tail -3 $O/r2_gpu_tests.log
for wl in ${WLS:-c0 c1 c2 c3 c5}; do
    timeout 300 python bench.py --workload $wl --steps 25 --no-cpu --e2e-steps 3 --no-configs --sustain-ms ${SUSTAIN:-0} > $O/r2_${T}_${wl}.json 2> $O/r2_${T}_${wl}.err
    python - <<PY
import json
try:
    d=json.loads(open("$O/r2_${T}_${wl}.json").read().strip().splitlines()[-1])
    s=d.get("sustained") or {}
    print("$wl", "us/frame %.1f" % d["roofline"]["launch_us"], "frac %.3f" % d["roofline"]["frac"], "sustained us %s frac %s" % (s.get("launch_us"), s.get("frac")), "same", d["device_equals_host_result"], d["clocks"]["sm_mhz"], d["clocks"]["reasons"], "pcie", d["e2e"].get("pcie_frac"))
except Exception as e:
    print("$wl failed", e)
PY
done
if [ "${FULL:-0}" = 1 ]; then
    timeout 900 python bench.py > $O/r2_${T}_bench.json 2> $O/r2_${T}_bench.err; echo "bench rc=$?"
    tail -c 2500 $O/r2_${T}_bench.json; tail -3 $O/r2_${T}_bench.err
fi
