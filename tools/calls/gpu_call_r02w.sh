#!/bin/bash
# 1 GPU: leaf variant with the next row group's first source row prefetched (-DNBODY_LEAF_PRE0=1): parity subset, bench A/B.
mkdir -p gpurun_out
LIB=$PWD/nbody_b200/libnbody_cuda_pre0.so
NBODY_CUDA_LIB=$LIB timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "bit_exact or expansions" > gpurun_out/r02w_parity.log 2>&1; echo "parity rc=$?"; tail -2 gpurun_out/r02w_parity.log
for tag in pre0 default pre0; do
	L=$PWD/nbody_b200/libnbody_cuda.so; [ $tag = pre0 ] && L=$LIB
	NBODY_CUDA_LIB=$L timeout 300 python bench.py --no-cpu-baseline --no-reference-capacity --no-config1 --no-accuracy --e2e-steps 1 > gpurun_out/r02w_bench_$tag.json 2>/dev/null
	python - <<PY
import json
d = json.load(open("gpurun_out/r02w_bench_$tag.json")); print("$tag", round(d["ms_per_step"], 3), "leaf", round(d["stage_ms"]["ms_leaf"], 2), "frac", round(d["p2p_fp32_tflops"]["tree_p2p_frac_of_peak"], 4))
PY
done
