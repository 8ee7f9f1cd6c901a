#!/bin/bash
# One gpurun call, 8 GPUs (round 2, after r02a / r02b have passed): strong scaling of the 16M Plummer step with and without the
# distributed sort, state check first. Optional first argument: a library variant (NBODY_CUDA_LIB) that r02a showed to be faster.
mkdir -p gpurun_out
[ -n "$1" ] && export NBODY_CUDA_LIB=$PWD/nbody_b200/libnbody_cuda_$1.so && echo "library: $NBODY_CUDA_LIB"
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
timeout 200 $TR --master-port 29655 tools/mg_check.py 2000000 4 plummer 0 64 > gpurun_out/r02c_mg_check_8gpu_dist_sort.log 2>&1; echo "rc=$?" >> gpurun_out/r02c_mg_check_8gpu_dist_sort.log
grep -h "MG_CHECK\|single-GPU vs\|state identical\|rc=" gpurun_out/r02c_mg_check_8gpu_dist_sort.log | cut -c1-320
for flags in 0 64; do
	timeout 200 $TR --master-port 2966$((flags / 64)) bench.py --gpus 8 --steps 8 --warmup 3 --no-reference-capacity --e2e-steps 2 --flags $flags \
		> gpurun_out/r02c_bench_16M_8gpu_flags$flags.json 2> gpurun_out/r02c_bench_8gpu_flags$flags.err; echo "bench flags=$flags rc=$?"
done
python - <<'PY'
import json
for f in ("r02c_bench_16M_8gpu_flags0.json", "r02c_bench_16M_8gpu_flags64.json"):
    try:
        d = json.load(open("gpurun_out/" + f)); print(f, round(d["ms_per_step"], 3), round(d["device_ms_per_step"], 3), {k: round(v, 2) for k, v in d["stage_ms"].items()}, round(d["e2e"]["ms_per_step"], 2))
    except Exception as e:
        print(f, "unreadable", e)
PY
