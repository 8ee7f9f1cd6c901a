#!/bin/bash
# 2 GPUs: the tree all-gather (plain ncclAllGather into equal slots) and the traversal-side marks with one process per GPU.
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 300 $TR --master-port 29711 bench.py --gpus 2 --no-cpu-baseline --no-reference-capacity --e2e-steps 2 --steps 6 --warmup 3 > gpurun_out/r02m_part16M_2gpu.json 2> gpurun_out/r02m.err; echo "rc=$?"
grep -v "OMP_NUM_THREADS\|^\*\*\*\*\|NCCL version" gpurun_out/r02m.err | tail -4 | cut -c1-400
python - <<'PY'
import json
d = json.load(open("gpurun_out/r02m_part16M_2gpu.json"))
print("ms/step", round(d["ms_per_step"], 3), {k[3:]: round(v, 2) for k, v in d["stage_ms"].items() if v}, "e2e ms", round(d["e2e"]["ms_per_step"], 2))
print("accuracy", d["accuracy"]["rms_rel"], "check", d["multi_gpu_check"]["pass"], d["multi_gpu_check"]["max_abs_dx"], "clocks", d["clocks"], "binding", d["config"]["cpu_binding"])
PY
