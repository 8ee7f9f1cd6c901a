#!/bin/bash
# 1 GPU: first run of the partitioned scheme (virtual ranks), the whole GPU suite with the new defaults, bench lines of the leaf variants.
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_partitioned.py -x -q -m gpu > gpurun_out/r02c_partitioned.log 2>&1; echo "partitioned rc=$?"; tail -40 gpurun_out/r02c_partitioned.log
timeout 900 python -m pytest tests -x -q -m gpu --deselect tests/test_gpu_partitioned.py > gpurun_out/r02c_gpu_suite.log 2>&1; echo "gpu suite rc=$?"; tail -5 gpurun_out/r02c_gpu_suite.log
for tag in default cta3 cta3_rows8 cta5_g8; do
	LIB=$PWD/nbody_b200/libnbody_cuda_$tag.so; [ $tag = default ] && LIB=$PWD/nbody_b200/libnbody_cuda.so
	NBODY_CUDA_LIB=$LIB timeout 200 python bench.py --no-cpu-baseline --no-reference-capacity --e2e-steps 1 > gpurun_out/r02c_bench_$tag.json 2> gpurun_out/r02c_bench_$tag.err; echo "bench $tag rc=$?"
	python - <<PY
import json
try:
    d = json.load(open("gpurun_out/r02c_bench_$tag.json")); print("$tag", round(d["ms_per_step"], 3), {k: round(v, 2) for k, v in d["stage_ms"].items()}, "leaf frac", round(d["p2p_fp32_tflops"]["tree_p2p_frac_of_peak"], 4))
except Exception as e:
    print("$tag unreadable", e)
PY
done
