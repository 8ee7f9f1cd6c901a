#!/bin/bash
# One gpurun call, 1 GPU: tree-build rewrite (parallel tile-count scan, 8 lanes per splitting node): parity first, then the bench line.
mkdir -p gpurun_out
timeout 200 python -m pytest tests/test_golden_fmm.py tests/test_gpu_parity.py -q -m gpu -x > gpurun_out/r01n_parity.log 2>&1; echo "rc=$?" >> gpurun_out/r01n_parity.log
tail -6 gpurun_out/r01n_parity.log
timeout 120 python bench.py --no-cpu-baseline > gpurun_out/r01n_bench_16M_1gpu.json 2> gpurun_out/r01n_bench.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.load(open('gpurun_out/r01n_bench_16M_1gpu.json'))
print(d['ms_per_step'], d['stage_ms'], d['reference_capacity']['stage_ms'])
PY
