#!/bin/bash
# ncu capture of the stream kernel on C0 (run under gpurun); $1 = tag
T=${1:-a}
O=gpurun_out
CMD="python bench.py --workload ${2:-c0} --steps 2 --warmup 3 --no-cpu --e2e-steps 1 --frames-per-step 2 --no-configs --sustain-ms 0"
timeout 300 $CMD > $O/r2_plain_$T.log 2>&1 || { echo "plain bench failed"; tail -5 $O/r2_plain_$T.log; exit 1; }
timeout 600 ncu --set full --clock-control none --import-source on -k regex:resize_stream_kernel -c 1 \
    --launch-skip 4 -f -o $O/r2_stream_$T $CMD > $O/r2_ncu_$T.log 2>&1
ls -la $O/r2_stream_$T.ncu-rep
