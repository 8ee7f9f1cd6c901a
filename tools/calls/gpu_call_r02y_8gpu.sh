#!/bin/bash
# 8 GPUs: the driver's command for N = 8 with the final code.
mkdir -p gpurun_out
timeout 280 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29702 bench.py --gpus 8 --steps 5 --warmup 3 > gpurun_out/r02y_part16M_8gpu.json 2> gpurun_out/r02y.err; echo "rc=$?"
grep -v "OMP_NUM_THREADS\|^\*\*\*\*\|NCCL version" gpurun_out/r02y.err | tail -4 | cut -c1-300
python - <<'PY'
import json
d = json.load(open("gpurun_out/r02y_part16M_8gpu.json"))
print("value", f'{d["value"]:.4g}', "ms/step", round(d["ms_per_step"], 3), {k[3:]: round(v, 2) for k, v in d["stage_ms"].items() if v}, "e2e ms", round(d["e2e"]["ms_per_step"], 2))
print("accuracy", d["accuracy"]["rms_rel"], "check", d["multi_gpu_check"]["pass"], "clocks", d["clocks"]["sm_mhz_min_over_ranks"], d["clocks"]["samples_in_timed_region"], "ref cap", d["reference_capacity"]["ms_per_step"])
PY
