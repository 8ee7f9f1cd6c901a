#!/bin/bash
# 1 GPU: partitioned scheme on virtual ranks (all cases, no -x), then the bench line with the accuracy check and config 1, reference arm.
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_partitioned.py -q -m gpu > gpurun_out/r02d_partitioned.log 2>&1; echo "partitioned rc=$?"; tail -60 gpurun_out/r02d_partitioned.log | cut -c1-300
timeout 400 python bench.py --no-reference-capacity > gpurun_out/r02d_bench.json 2> gpurun_out/r02d_bench.err; echo "bench rc=$?"; tail -3 gpurun_out/r02d_bench.err; cut -c1-1500 gpurun_out/r02d_bench.json
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r02d_bench_reference.json 2> gpurun_out/r02d_bench_reference.err; echo "reference rc=$?"; cut -c1-1200 gpurun_out/r02d_bench_reference.json
