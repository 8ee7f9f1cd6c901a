#!/bin/bash
# 2 GPUs: first run of the partitioned scheme with one process per GPU (cudaIpc peer pointers + NCCL), small N first.
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
run() {  # tag, port, extra args
	timeout 300 $TR --master-port $2 bench.py --gpus 2 --no-cpu-baseline --no-reference-capacity --e2e-steps 2 ${@:3} > gpurun_out/r02f_$1.json 2> gpurun_out/r02f_$1.err; echo "$1 rc=$?"
	tail -4 gpurun_out/r02f_$1.err | cut -c1-400
	python - <<PY
import json
try:
    d = json.load(open("gpurun_out/r02f_$1.json"))
    print("$1", "ms/step", round(d["ms_per_step"], 3), "dev", round(d["device_ms_per_step"], 3), {k: round(v, 2) for k, v in d["stage_ms"].items()}, "e2e ms", round(d["e2e"]["ms_per_step"], 2))
    print("   accuracy", d.get("accuracy"), "\n   check", d.get("multi_gpu_check"), "\n   per rank", d.get("per_rank_ms"))
    c = d["counts"]; print("   counts", {k: c[k] for k in ("n_particles", "n_nodes", "halo_particles", "imported_nodes", "migrated_particles", "device_bytes", "work_imbalance", "retries") if k in c})
except Exception as e:
    print("$1 unreadable", e)
PY
}
run small 29701 --n 2000000 --steps 4 --warmup 3
run part16M 29702 --steps 6 --warmup 3
run repl16M 29703 --steps 6 --warmup 3 --scheme replicated --no-accuracy --no-multi-check
run repl16M_distsort 29704 --steps 6 --warmup 3 --scheme replicated --flags 64 --no-accuracy
