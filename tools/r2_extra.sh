#!/bin/bash
O=gpurun_out
timeout 1800 python -m pytest tests -x -q -m gpu > $O/r2_gpu_tests.log 2>&1; echo "gpu tests rc=$?"; tail -3 $O/r2_gpu_tests.log
timeout 300 python tools/hicomp_probe.py 2>&1 | tail -8
timeout 300 python tools/nonfinite_probe.py 2>&1 | tail -8
