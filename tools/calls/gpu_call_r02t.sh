#!/bin/bash
# 1 GPU: ncu --set full of the final leaf kernel (chunk tickets, TMA fill, uniform dispatch, async entry prefetch), CSV exports.
mkdir -p gpurun_out
timeout 500 ncu --set full --clock-control none --import-source on -k 'regex:^k_leaf$' -s 1 -c 1 -o /tmp/r02t_k_leaf python tools/prof_step.py 16777216 1 4 48 > gpurun_out/r02t_ncu.log 2>&1; echo "ncu rc=$?"
ncu -i /tmp/r02t_k_leaf.ncu-rep --page raw --csv > gpurun_out/r02t_k_leaf_raw.csv 2>/dev/null
ncu -i /tmp/r02t_k_leaf.ncu-rep --page source --csv > gpurun_out/r02t_k_leaf_source.csv 2>/dev/null
python tools/ncu_summary.py /tmp/r02t_k_leaf.ncu-rep k_leaf 40 > gpurun_out/r02t_k_leaf_summary.txt 2>&1; head -60 gpurun_out/r02t_k_leaf_summary.txt | cut -c1-170
