#!/bin/bash
# One gpurun call, 2 GPUs (round 2): the distributed sort (NBODY_FLAG_DIST_SORT = 64: slice sort + all-gather of the runs + pairwise
# merges, written in round 1 after the GPU budget was spent; the merge arithmetic is CPU-tested by tests/test_merge_host.py, the
# kernel and the in-step collective have NOT run on hardware). State check against the 1-GPU run first, then A/B bench lines.
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 150 $TR --master-port 29655 tools/mg_check.py 200000 5 plummer 0 64 > gpurun_out/r02b_mg_check_dist_sort.log 2>&1; echo "rc=$?" >> gpurun_out/r02b_mg_check_dist_sort.log
grep -h "MG_CHECK\|single-GPU vs\|state identical\|rc=" gpurun_out/r02b_mg_check_dist_sort.log | cut -c1-320
timeout 150 $TR --master-port 29656 tools/mg_check.py 3000000 4 plummer 0 64 > gpurun_out/r02b_mg_check_dist_sort_3M.log 2>&1; echo "rc=$?" >> gpurun_out/r02b_mg_check_dist_sort_3M.log
grep -h "MG_CHECK\|single-GPU vs\|state identical\|rc=" gpurun_out/r02b_mg_check_dist_sort_3M.log | cut -c1-320
timeout 150 $TR --master-port 29657 bench.py --gpus 2 --steps 6 --warmup 3 --no-reference-capacity --e2e-steps 2 > gpurun_out/r02b_bench_16M_2gpu.json 2> gpurun_out/r02b_bench_2gpu.err; echo "bench rc=$?"
timeout 150 $TR --master-port 29658 bench.py --gpus 2 --steps 6 --warmup 3 --no-reference-capacity --e2e-steps 2 --flags 64 > gpurun_out/r02b_bench_16M_2gpu_dist_sort.json 2> gpurun_out/r02b_bench_2gpu_dist_sort.err; echo "bench rc=$?"
NBODY_NO_COMM_SPLIT=1 timeout 150 $TR --master-port 29659 bench.py --gpus 2 --steps 6 --warmup 3 --no-reference-capacity --e2e-steps 2 --flags 64 > gpurun_out/r02b_bench_16M_2gpu_dist_sort_one_comm.json 2> gpurun_out/r02b_bench_2gpu_dist_sort_one_comm.err; echo "bench rc=$?"
python - <<'PY'
import json
for f in ("r02b_bench_16M_2gpu.json", "r02b_bench_16M_2gpu_dist_sort.json", "r02b_bench_16M_2gpu_dist_sort_one_comm.json"):
    try:
        d = json.load(open("gpurun_out/" + f)); print(f, round(d["ms_per_step"], 3), round(d["device_ms_per_step"], 3), {k: round(v, 2) for k, v in d["stage_ms"].items()}, round(d["e2e"]["ms_per_step"], 2))
    except Exception as e:
        print(f, "unreadable", e)
PY
