#!/bin/bash
# ncu evidence for profiles/: run on the GPU box via gpurun.  Every profiled
# command is first run plain and must exit 0.
set -u
O=gpurun_out
CMD="python bench.py --steps 2 --warmup 3 --no-cpu --e2e-steps 1 --frames-per-step 2"
$CMD > $O/plain_v2.log 2>&1 || { echo "plain bench failed"; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv \
    --log-file $O/launches_r1_v2.csv $CMD > $O/ncu_launch_v2.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:resize_fused_kernel -c 1 \
    --launch-skip 4 -f -o $O/c0_final2 $CMD > $O/ncu_c0.log 2>&1
python tools/bench_extra.py warp > $O/warp_plain.log 2>&1 || { echo "warp failed"; exit 1; }
ncu --set full --clock-control none --import-source on -k regex:warp_kernel -c 1 \
    -f -o $O/warp_r1 python tools/bench_extra.py warp > $O/ncu_warp.log 2>&1
python tools/bench_extra.py conv > $O/conv_plain.log 2>&1 || { echo "conv failed"; exit 1; }
ncu --set full --clock-control none --import-source on -k regex:convolve_kernel -c 1 \
    -f -o $O/conv_r1 python tools/bench_extra.py conv > $O/ncu_conv.log 2>&1
ls -la $O/*.ncu-rep | tail -3
tail -2 $O/plain_v2.log | cut -c1-300
