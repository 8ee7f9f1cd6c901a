#!/bin/bash
# 1 GPU: M2L with the cross-item prefetch: parity (expansions, accelerations, partitioned), bench.
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_golden_fmm.py tests/test_gpu_parity.py tests/test_gpu_partitioned.py -x -q -m gpu > gpurun_out/r02p_parity.log 2>&1; echo "parity rc=$?"; tail -4 gpurun_out/r02p_parity.log | cut -c1-300
for i in 1 2; do
timeout 300 python bench.py --no-cpu-baseline --no-reference-capacity --no-config1 --e2e-steps 1 > gpurun_out/r02p_bench.json 2> gpurun_out/r02p_bench.err; echo "bench rc=$?"
python - <<'PY'
import json
d = json.load(open("gpurun_out/r02p_bench.json")); print(round(d["ms_per_step"], 3), {k: round(v, 2) for k, v in d["stage_ms"].items() if v}, "m2l TF", round(d["m2l_fp32_tflops"], 2), "accuracy", d["accuracy"]["rms_rel"])
PY
done
timeout 300 python tools/sweep.py 16777216 "cap=8;steps=2" > gpurun_out/r02p_cap8.log 2>&1; cat gpurun_out/r02p_cap8.log | cut -c1-300
