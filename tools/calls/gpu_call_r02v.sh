#!/bin/bash
# 1 GPU: the flat-entry leaf variant (tiles across segment ends, -DNBODY_LEAF_FLAT=1) against the default: parity, bench, ncu raw metrics.
mkdir -p gpurun_out
LIB=$PWD/nbody_b200/libnbody_cuda_flat.so
NBODY_CUDA_LIB=$LIB timeout 900 python -m pytest tests/test_golden_fmm.py tests/test_gpu_parity.py -x -q -m gpu -k "not full_size and not distributed_sort" > gpurun_out/r02v_parity_flat.log 2>&1; echo "parity flat rc=$?"; tail -2 gpurun_out/r02v_parity_flat.log
for tag in flat default flat default; do
	L=$PWD/nbody_b200/libnbody_cuda.so; [ $tag = flat ] && L=$LIB
	NBODY_CUDA_LIB=$L timeout 300 python bench.py --no-cpu-baseline --no-reference-capacity --no-config1 --no-accuracy --e2e-steps 1 > gpurun_out/r02v_bench_$tag.json 2>/dev/null
	python - <<PY
import json
d = json.load(open("gpurun_out/r02v_bench_$tag.json")); print("$tag", round(d["ms_per_step"], 3), "leaf", round(d["stage_ms"]["ms_leaf"], 2), "frac", round(d["p2p_fp32_tflops"]["tree_p2p_frac_of_peak"], 4))
PY
done
NBODY_CUDA_LIB=$LIB timeout 400 ncu --set full --clock-control none --import-source on -k 'regex:^k_leaf$' -s 1 -c 1 -o /tmp/r02v_flat python tools/prof_step.py 16777216 1 4 48 > gpurun_out/r02v_ncu.log 2>&1; echo "ncu rc=$?"
python tools/ncu_summary.py /tmp/r02v_flat.ncu-rep k_leaf 14 > gpurun_out/r02v_k_leaf_flat_summary.txt 2>&1; head -34 gpurun_out/r02v_k_leaf_flat_summary.txt | cut -c1-170
