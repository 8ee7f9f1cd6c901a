#!/bin/bash
# One gpurun call, 1 GPU: the repaired driver test, then configs 4 and 5 at their per-GPU particle counts on a single B200.
mkdir -p gpurun_out
timeout 120 python -m pytest tests/test_gpu_cpp_and_multi.py -q -m gpu > gpurun_out/r01m_cpp_tests.log 2>&1; echo "rc=$?" >> gpurun_out/r01m_cpp_tests.log
tail -4 gpurun_out/r01m_cpp_tests.log
timeout 200 python tests/tools/scale_evidence.py two_galaxies 67108864 48 > gpurun_out/r01m_scale_two_galaxies_64M.log 2>&1; echo "rc=$?" >> gpurun_out/r01m_scale_two_galaxies_64M.log
cut -c1-700 gpurun_out/r01m_scale_two_galaxies_64M.log | tail -8
timeout 240 python tests/tools/scale_evidence.py uniform 134217728 80 0.6 2 > gpurun_out/r01m_scale_uniform_128M.log 2>&1; echo "rc=$?" >> gpurun_out/r01m_scale_uniform_128M.log
cut -c1-700 gpurun_out/r01m_scale_uniform_128M.log | tail -8
