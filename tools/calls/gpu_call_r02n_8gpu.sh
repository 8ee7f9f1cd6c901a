#!/bin/bash
# 8 GPUs: the driver's own command for N = 8 (default bench line), then config 4 with the final code.
mkdir -p gpurun_out
run() {  # tag, ranks, port, extra args
	local TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $2 --master-addr 127.0.0.1"
	timeout ${TMO:-300} $TR --master-port $3 bench.py --gpus $2 ${@:4} > gpurun_out/r02n_$1.json 2> gpurun_out/r02n_$1.err; echo "$1 rc=$?"
	grep -v "OMP_NUM_THREADS\|^\*\*\*\*\|NCCL version" gpurun_out/r02n_$1.err | tail -4 | cut -c1-400
	python - <<PY
import json
try:
    d = json.load(open("gpurun_out/r02n_$1.json"))
    print("$1", "value", f'{d["value"]:.4g}', "ms/step", round(d["ms_per_step"], 3), "dev", round(d["device_ms_per_step"], 3), {k[3:]: round(v, 2) for k, v in d["stage_ms"].items() if v}, "e2e ms", round(d["e2e"]["ms_per_step"], 2))
    print("   accuracy", (d.get("accuracy") or {}).get("rms_rel"), "check", (d.get("multi_gpu_check") or {}).get("pass"), "clocks", {k: v for k, v in d["clocks"].items() if k != "per_rank"}, [c["sm_mhz"] for c in d["clocks"].get("per_rank", [])], d["config"].get("cpu_binding"))
    print("   per rank", d["per_rank_ms"]["columns"]); [print("     ", r) for r in d["per_rank_ms"]["rows"]]
    c = d["counts"]; print("   counts", {k: c[k] for k in ("n_particles", "n_nodes", "halo_particles", "imported_nodes", "migrated_particles", "device_bytes", "retries", "work_imbalance_per_step") if k in c})
except Exception as e:
    print("$1 unreadable", e)
PY
}
run part16M_default 8 29702 --steps 5 --warmup 3
run part16M_4gpu 4 29703 --steps 5 --warmup 3 --no-accuracy
TMO=500 run config4_two_galaxies_64M 8 29705 --workload two_galaxies --particles 67108864 --steps 6 --warmup 3 --no-multi-check --accuracy-targets 16384 --no-reference-capacity
