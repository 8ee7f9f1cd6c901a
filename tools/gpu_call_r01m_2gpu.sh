#!/bin/bash
# One gpurun call, 2 GPUs: distributed step == single-GPU step (fixed and variable time step), then a short 2-GPU bench line.
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 120 $TR --master-port 29655 tools/mg_check.py 200000 4 > gpurun_out/r01m_mg_check.log 2>&1; echo "rc=$?" >> gpurun_out/r01m_mg_check.log
timeout 120 $TR --master-port 29656 tools/mg_check.py 100000 4 plummer 0.05 > gpurun_out/r01m_mg_check_eta.log 2>&1; echo "rc=$?" >> gpurun_out/r01m_mg_check_eta.log
grep -h "MG_CHECK\|time steps\|single-GPU vs\|imbalance\|rc=" gpurun_out/r01m_mg_check.log gpurun_out/r01m_mg_check_eta.log | cut -c1-300
timeout 150 $TR --master-port 29657 bench.py --gpus 2 --steps 5 --warmup 3 --no-reference-capacity > gpurun_out/r01m_bench_16M_2gpu.json 2> gpurun_out/r01m_bench_2gpu.err
echo "bench rc=$?"; cut -c1-300 gpurun_out/r01m_bench_16M_2gpu.json; tail -3 gpurun_out/r01m_bench_2gpu.err
