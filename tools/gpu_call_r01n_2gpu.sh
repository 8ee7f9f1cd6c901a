#!/bin/bash
# One gpurun call, 2 GPUs: velocity exchange on a second stream (overlapped with the next step): state check, then A/B bench lines.
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 100 $TR --master-port 29655 tools/mg_check.py 200000 5 > gpurun_out/r01n_mg_check.log 2>&1; echo "rc=$?" >> gpurun_out/r01n_mg_check.log
grep -h "MG_CHECK\|single-GPU vs\|state identical\|rc=" gpurun_out/r01n_mg_check.log | cut -c1-320
timeout 100 $TR --master-port 29656 tools/mg_check.py 100000 4 plummer 0.05 > gpurun_out/r01n_mg_check_eta.log 2>&1; echo "rc=$?" >> gpurun_out/r01n_mg_check_eta.log
grep -h "MG_CHECK\|time steps\|rc=" gpurun_out/r01n_mg_check_eta.log | cut -c1-200
timeout 100 $TR --master-port 29657 bench.py --gpus 2 --steps 6 --warmup 3 --no-reference-capacity --e2e-steps 2 > gpurun_out/r01n_bench_16M_2gpu.json 2> gpurun_out/r01n_bench_2gpu.err; echo "bench rc=$?"
timeout 100 $TR --master-port 29658 bench.py --gpus 2 --steps 6 --warmup 3 --no-reference-capacity --e2e-steps 2 --flags 32 > gpurun_out/r01n_bench_16M_2gpu_no_overlap.json 2> gpurun_out/r01n_bench_2gpu_no_overlap.err; echo "bench rc=$?"
python - <<'PY'
import json
for f in ("r01n_bench_16M_2gpu.json", "r01n_bench_16M_2gpu_no_overlap.json"):
    try:
        d = json.load(open("gpurun_out/" + f)); print(f, round(d["ms_per_step"], 3), round(d["device_ms_per_step"], 3), {k: round(v, 2) for k, v in d["stage_ms"].items()}, round(d["e2e"]["ms_per_step"], 2))
    except Exception as e:
        print(f, "unreadable", e)
PY
