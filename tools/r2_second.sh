#!/bin/bash
O=gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu > $O/r2_parity.log 2>&1; echo "parity rc=$?"
tail -3 $O/r2_parity.log
bash tools/r2_prof.sh a c0
