#!/bin/bash
# 8 GPUs: 16M Plummer with the multipole pull (A/B of the leaf hand-out order), then BASELINE config 5: weak scaling of the uniform
# cube at 2^27 particles per GPU (2 / 4 / 8 ranks here; 1 rank in a 1-GPU call).
mkdir -p gpurun_out
free -g | head -2; nvidia-smi --query-gpu=memory.total --format=csv,noheader | head -1
run() {  # tag, ranks, port, extra args
	local TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $2 --master-addr 127.0.0.1"
	timeout ${TMO:-300} $TR --master-port $3 bench.py --gpus $2 --no-cpu-baseline --no-reference-capacity ${@:4} > gpurun_out/r02i_$1.json 2> gpurun_out/r02i_$1.err; echo "$1 rc=$?"
	grep -v "OMP_NUM_THREADS\|^\*\*\*\*\|NCCL version" gpurun_out/r02i_$1.err | tail -4 | cut -c1-400
	python - <<PY
import json
try:
    d = json.load(open("gpurun_out/r02i_$1.json"))
    print("$1", "value", f'{d["value"]:.4g}', "ms/step", round(d["ms_per_step"], 3), "dev", round(d["device_ms_per_step"], 3), {k[3:]: round(v, 2) for k, v in d["stage_ms"].items()}, "e2e ms", round(d["e2e"]["ms_per_step"], 2))
    print("   accuracy", (d.get("accuracy") or {}).get("rms_rel"), "check", (d.get("multi_gpu_check") or {}).get("pass"))
    print("   per rank", d["per_rank_ms"]["columns"]); [print("     ", r) for r in d["per_rank_ms"]["rows"]]
    c = d["counts"]; print("   counts", {k: c[k] for k in ("n_particles", "n_nodes", "halo_particles", "imported_nodes", "migrated_particles", "device_bytes", "retries", "work_imbalance_per_step") if k in c})
except Exception as e:
    print("$1 unreadable", e)
PY
}
run part16M 8 29702 --steps 8 --warmup 3 --e2e-steps 2
NBODY_LEAF_REVERSE=0 run part16M_forward_leaf_order 8 29703 --steps 8 --warmup 3 --e2e-steps 1 --no-accuracy --no-multi-check
TMO=600 run config5_uniform_2p30_8gpu 8 29704 --workload uniform --particles 1073741824 --leaf-capacity 80 --steps 3 --warmup 3 --no-accuracy --no-multi-check
TMO=600 run config5_uniform_2p29_4gpu 4 29705 --workload uniform --particles 536870912 --leaf-capacity 80 --steps 3 --warmup 3 --no-accuracy --no-multi-check
TMO=700 run config5_uniform_2p28_2gpu 2 29706 --workload uniform --particles 268435456 --leaf-capacity 80 --steps 3 --warmup 3 --accuracy-targets 8192 --no-multi-check
