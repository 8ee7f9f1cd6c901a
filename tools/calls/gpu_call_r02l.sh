#!/bin/bash
# 1 GPU: GPU suite with the fused marks / all-gather layout (virtual ranks), bench; (the NCCL side is checked on 2 GPUs next)
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/r02l_gpu_suite.log 2>&1; echo "gpu suite rc=$?"; tail -5 gpurun_out/r02l_gpu_suite.log | cut -c1-300
timeout 400 python bench.py --no-reference-capacity > gpurun_out/r02l_bench.json 2> gpurun_out/r02l_bench.err; echo "bench rc=$?"; tail -2 gpurun_out/r02l_bench.err
python - <<'PY'
import json
d = json.load(open("gpurun_out/r02l_bench.json")); print(round(d["ms_per_step"], 3), {k: round(v, 2) for k, v in d["stage_ms"].items() if v}, "leaf frac", round(d["p2p_fp32_tflops"]["tree_p2p_frac_of_peak"], 4), d["accuracy"]["rms_rel"], d["clocks"], d["gpu_launches"])
PY
timeout 300 python tools/virt_prof.py 16777216 8 4 > gpurun_out/r02l_virt8.log 2>&1; tail -1 gpurun_out/r02l_virt8.log | cut -c1-400
