#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_partitioned.py -q -m gpu > gpurun_out/r02e_partitioned.log 2>&1; echo "partitioned rc=$?"; tail -30 gpurun_out/r02e_partitioned.log | cut -c1-300
timeout 900 python -m pytest tests/test_golden_fmm.py tests/test_gpu_parity.py -x -q -m gpu > gpurun_out/r02e_parity.log 2>&1; echo "parity rc=$?"; tail -3 gpurun_out/r02e_parity.log
timeout 300 python bench.py --no-cpu-baseline --no-reference-capacity --no-config1 --e2e-steps 1 > gpurun_out/r02e_bench.json 2> gpurun_out/r02e_bench.err; echo "bench rc=$?"
python - <<'PY'
import json
d = json.load(open("gpurun_out/r02e_bench.json")); print(round(d["ms_per_step"], 3), {k: round(v, 2) for k, v in d["stage_ms"].items()}, "leaf frac", round(d["p2p_fp32_tflops"]["tree_p2p_frac_of_peak"], 4), d["accuracy"]["rms_rel"])
PY
