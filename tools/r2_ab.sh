#!/bin/bash
# A/B: old vs new stream kernel (B2R_LIBPATH) x band search off / on
O=gpurun_out
for wl in ${WLS:-c0 c2 c5 c1}; do
  for lib in base new; do
    for bs in ${BSS:-0 1}; do
      L=openimageio_b200/libb2r.so; [ $lib = base ] && L=openimageio_b200/libb2r_base.so
      B2R_LIBPATH=$PWD/$L B2R_BANDSEARCH=$bs timeout 300 python bench.py --workload $wl --steps 25 --no-cpu --e2e-steps 3 --no-configs --sustain-ms 0 ${EXTRA} > $O/r2_ab_${wl}_${lib}_$bs.json 2> $O/r2_ab_${wl}_${lib}_$bs.err
      python - <<PY
import json
try:
    d=json.loads(open("$O/r2_ab_${wl}_${lib}_$bs.json").read().strip().splitlines()[-1])
    s=d.get("sustained") or {}
    print("$wl lib=$lib search=$bs", "us/frame %.1f" % d["roofline"]["launch_us"], "frac %.3f" % d["roofline"]["frac"], "sustained %s %s" % (s.get("launch_us"), (s.get("clocks") or {}).get("sm_mhz")), "same", d["device_equals_host_result"], d["clocks"]["sm_mhz"], d["clocks"]["reasons"])
except Exception as e:
    print("$wl $lib $bs failed", e)
PY
    done
  done
done
