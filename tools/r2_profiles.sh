#!/bin/bash
# round-2 ncu evidence for profiles/ (run under gpurun): launch list of the
# bench command + one --set full capture of the dominant kernel per workload.
# Every profiled command is first run plain and must exit 0.
O=gpurun_out
BASE="--steps 2 --warmup 3 --no-cpu --e2e-steps 1 --frames-per-step 2 --no-configs --sustain-ms 0"
timeout 300 python bench.py --workload c0 $BASE > $O/r2p_plain_c0.log 2>&1 || { echo "plain c0 failed"; tail -5 $O/r2p_plain_c0.log; exit 1; }
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv \
    --log-file $O/launches_r2.csv python bench.py --workload c0 $BASE > $O/r2p_ncu_launch.log 2>&1
for wl in c0 c2 c5 c1 c3; do
    K=resize_stream_kernel
    [ $wl = c3 ] && K=resize_expand_kernel
    timeout 300 python bench.py --workload $wl $BASE > $O/r2p_plain_$wl.log 2>&1 || { echo "plain $wl failed"; continue; }
    timeout 600 ncu --set full --clock-control none --import-source on -k regex:$K -c 1 \
        --launch-skip 4 -f -o $O/r2p_$wl python bench.py --workload $wl $BASE > $O/r2p_ncu_$wl.log 2>&1
    ls -la $O/r2p_$wl.ncu-rep
done
