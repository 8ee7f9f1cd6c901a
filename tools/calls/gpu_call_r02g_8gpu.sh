#!/bin/bash
# 8 GPUs: strong scaling of the 16M Plummer step, partitioned (own particles + LET) against replicated + distributed sort; then
# BASELINE config 4 (two-galaxy 2^26, per-step rebalancing).
mkdir -p gpurun_out
W=${W:-8}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $W --master-addr 127.0.0.1"
run() {  # tag, port, extra args
	timeout ${TMO:-300} $TR --master-port $2 bench.py --gpus $W --no-cpu-baseline --no-reference-capacity --e2e-steps 2 ${@:3} > gpurun_out/r02g_$1.json 2> gpurun_out/r02g_$1.err; echo "$1 rc=$?"
	grep -v "OMP_NUM_THREADS\|^\*\*\*\*\|NCCL version" gpurun_out/r02g_$1.err | tail -4 | cut -c1-400
	python - <<PY
import json
try:
    d = json.load(open("gpurun_out/r02g_$1.json"))
    print("$1", "ms/step", round(d["ms_per_step"], 3), "dev", round(d["device_ms_per_step"], 3), {k[3:]: round(v, 2) for k, v in d["stage_ms"].items()}, "e2e ms", round(d["e2e"]["ms_per_step"], 2))
    print("   accuracy", (d.get("accuracy") or {}).get("rms_rel"), "check", {k: v for k, v in (d.get("multi_gpu_check") or {}).items() if k in ("pass", "max_abs_dx", "first_step_same_tree_order_as_1gpu", "device_bytes_per_rank")})
    print("   per rank", d["per_rank_ms"]["columns"]); [print("     ", r) for r in d["per_rank_ms"]["rows"]]
    c = d["counts"]; print("   counts", {k: c[k] for k in ("n_particles", "n_nodes", "halo_particles", "imported_nodes", "migrated_particles", "device_bytes", "retries", "work_imbalance_per_step") if k in c})
except Exception as e:
    print("$1 unreadable", e)
PY
}
run part16M 29702 --steps 8 --warmup 3
run repl16M_distsort 29704 --steps 8 --warmup 3 --scheme replicated --flags 64 --no-accuracy --no-multi-check
TMO=500 run config4_two_galaxies_64M 29705 --workload two_galaxies --particles 67108864 --steps 6 --warmup 3 --no-multi-check --accuracy-targets 16384
