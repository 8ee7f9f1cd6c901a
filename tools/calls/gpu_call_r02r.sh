#!/bin/bash
# 1 GPU: expansion order 5 (parity against the oracle), whole parity file, order-5 timing at 2^24.
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_parity.py tests/test_golden_fmm.py -x -q -m gpu > gpurun_out/r02r_parity.log 2>&1; echo "parity rc=$?"; tail -6 gpurun_out/r02r_parity.log | cut -c1-300
timeout 400 python tools/sweep.py 16777216 "cap=48;order=5" "cap=48;order=5;tau=0" "cap=48;order=4" > gpurun_out/r02r_order5.log 2>&1; cat gpurun_out/r02r_order5.log | cut -c1-330
timeout 300 python tests/tools/accuracy_full.py plummer 16777216 48 > gpurun_out/r02r_acc4.log 2>&1; tail -1 gpurun_out/r02r_acc4.log | cut -c1-300
