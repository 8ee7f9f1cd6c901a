#!/bin/bash
# One gpurun call, 1 GPU: new GPU tests first, then the default bench line, then the whole GPU suite if time remains.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm --format=csv,noheader > gpurun_out/r01m_gpu.txt 2>&1
timeout 200 python -m pytest tests/test_gpu_timestep_checkpoint.py tests/test_gpu_cpp_and_multi.py -q -m gpu > gpurun_out/r01m_new_tests.log 2>&1
echo "rc=$?" >> gpurun_out/r01m_new_tests.log
tail -15 gpurun_out/r01m_new_tests.log
timeout 170 python bench.py > gpurun_out/r01m_bench_16M_1gpu.json 2> gpurun_out/r01m_bench.err
echo "bench rc=$?"
cut -c1-400 gpurun_out/r01m_bench_16M_1gpu.json
timeout 240 python -m pytest tests -q -m gpu -x --deselect tests/test_gpu_timestep_checkpoint.py --deselect tests/test_gpu_cpp_and_multi.py > gpurun_out/r01m_all_gpu_tests.log 2>&1
echo "rc=$?" >> gpurun_out/r01m_all_gpu_tests.log
tail -5 gpurun_out/r01m_all_gpu_tests.log
