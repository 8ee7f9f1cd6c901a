#!/bin/bash
# 1 GPU: leaf kernel with the asynchronous entry prefetch: parity, bench (twice: the leaf kernel varies by ~0.5 ms between runs).
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_golden_fmm.py tests/test_gpu_parity.py tests/test_gpu_partitioned.py -x -q -m gpu > gpurun_out/r02s_parity.log 2>&1; echo "parity rc=$?"; tail -3 gpurun_out/r02s_parity.log | cut -c1-300
for i in 1 2 3; do
timeout 300 python bench.py --no-cpu-baseline --no-reference-capacity --no-config1 --no-accuracy --e2e-steps 1 > gpurun_out/r02s_bench.json 2> gpurun_out/r02s_bench.err; echo "bench rc=$?"
python - <<'PY'
import json
d = json.load(open("gpurun_out/r02s_bench.json")); print(round(d["ms_per_step"], 3), {k: round(v, 2) for k, v in d["stage_ms"].items() if v}, "leaf frac", round(d["p2p_fp32_tflops"]["tree_p2p_frac_of_peak"], 4))
PY
done
