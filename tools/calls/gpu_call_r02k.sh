#!/bin/bash
# 1 GPU: the leaf kernel on a rank's 1/8 share, through virtual ranks (16M Plummer, 8 ranks): the three work-item modes, ncu of mode 0.
mkdir -p gpurun_out
for mode in 0 1 2; do
	NBODY_LEAF_ITEMS=$mode timeout 300 python tools/virt_prof.py 16777216 8 4 > gpurun_out/r02k_virt8_mode$mode.log 2>&1; echo "mode $mode rc=$?"; tail -1 gpurun_out/r02k_virt8_mode$mode.log | cut -c1-400
done
for mode in 0 1 2; do
	NBODY_LEAF_ITEMS=$mode timeout 200 python bench.py --no-cpu-baseline --no-reference-capacity --no-config1 --no-accuracy --e2e-steps 1 > gpurun_out/r02k_bench_mode$mode.json 2>/dev/null
	python - <<PY
import json
d = json.load(open("gpurun_out/r02k_bench_mode$mode.json")); print("1 GPU mode $mode", round(d["ms_per_step"], 3), "leaf", round(d["stage_ms"]["ms_leaf"], 2), "frac", round(d["p2p_fp32_tflops"]["tree_p2p_frac_of_peak"], 4))
PY
done
NBODY_LEAF_ITEMS=0 timeout 600 ncu --set full --clock-control none -k 'regex:^k_leaf$' -s 16 -c 8 -o /tmp/r02k_leaf8 python tools/virt_prof.py 16777216 8 3 > gpurun_out/r02k_ncu.log 2>&1; echo "ncu rc=$?"
ncu -i /tmp/r02k_leaf8.ncu-rep --page raw --csv > gpurun_out/r02k_leaf8_raw.csv 2>/dev/null
python tools/ncu_summary.py /tmp/r02k_leaf8.ncu-rep > gpurun_out/r02k_leaf8_summary.txt 2>&1; head -20 gpurun_out/r02k_leaf8_summary.txt
