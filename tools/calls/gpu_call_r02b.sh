#!/bin/bash
# 1 GPU: ncu --set full of the leaf kernel (TMA tile fill, now the default), k_traverse and k_m2l, exported as CSV / text on the box
# (the .ncu-rep files stay in /tmp there: gpurun_out/ is capped at 64 MiB), then a capacity / tau sweep with the new fill.
mkdir -p gpurun_out
cap() {  # kernel regex, launches to skip, launches to capture
	timeout 500 ncu --set full --clock-control none --import-source on -k regex:$1 -s $2 -c $3 -o /tmp/r02b_$1 \
		python tools/prof_step.py 16777216 1 4 48 > gpurun_out/r02b_ncu_$1.log 2>&1; echo "ncu $1 rc=$?"
	ncu -i /tmp/r02b_$1.ncu-rep --page raw --csv > gpurun_out/r02b_$1_raw.csv 2>/dev/null
	python tools/ncu_summary.py /tmp/r02b_$1.ncu-rep $1 28 > gpurun_out/r02b_$1_summary.txt 2>&1
}
cap k_leaf 0 1
ncu -i /tmp/r02b_k_leaf.ncu-rep --page source --csv > gpurun_out/r02b_k_leaf_source.csv 2>/dev/null
cap k_m2l 0 2
cap k_traverse 21 21
ls -la gpurun_out/ | tail -12
timeout 600 python tools/sweep.py 16777216 "cap=32" "cap=40" "cap=48" "cap=56" "cap=64" "cap=80" "cap=48;tau=0.16" "cap=56;tau=0.16" > gpurun_out/r02b_sweep.log 2>&1; cat gpurun_out/r02b_sweep.log
