#!/bin/bash
# A/B: the in-tree library against openimageio_b200/libb2r_exp.so (built with an
# experiment macro), same box, same bench.  usage: ab_bench.sh "c2 c5 c0"
for wl in ${1:-c2}; do
  for lib in libb2r.so libb2r_exp.so; do
    B2R_LIBPATH=/root/repo/openimageio_b200/$lib python bench.py --workload $wl --steps 25 --no-cpu --e2e-steps 3 > gpurun_out/ab_${wl}_${lib}.json 2>&1
  done
done
