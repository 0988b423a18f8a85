#!/bin/bash
# full GPU test suite + the default bench line (all configs) + the reference arm
O=gpurun_out
timeout 2400 python -m pytest tests -x -q -m gpu > $O/r2_gpu_tests.log 2>&1; echo "gpu tests rc=$?"
tail -4 $O/r2_gpu_tests.log
timeout 900 python bench.py > $O/r2_bench.json 2> $O/r2_bench.err; echo "bench rc=$?"
tail -c 3000 $O/r2_bench.json
tail -5 $O/r2_bench.err
