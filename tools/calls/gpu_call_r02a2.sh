#!/bin/bash
# Second half of the first round-2 call (1 GPU): accuracy evidence at the benched configuration, ncu --set full of k_traverse / k_m2l,
# compute-sanitizer memcheck + racecheck at N = 30k.
mkdir -p gpurun_out
timeout 300 python tests/tools/accuracy_full.py plummer 16777216 48 > gpurun_out/r02a_accuracy_plummer16M_cap48.log 2>&1; echo "accuracy rc=$?"; tail -2 gpurun_out/r02a_accuracy_plummer16M_cap48.log
timeout 500 ncu --set full --clock-control none --import-source on -k regex:"k_traverse|k_m2l" -o gpurun_out/r02a_traverse_m2l \
	python tools/prof_step.py 16777216 1 4 48 > gpurun_out/r02a_ncu_traverse_m2l.log 2>&1; echo "ncu traverse/m2l rc=$?"
for tool in memcheck racecheck; do
	timeout 400 compute-sanitizer --tool $tool --error-exitcode 9 python tools/prof_step.py 30000 1 4 8 > gpurun_out/r02a_sanitizer_$tool.log 2>&1; echo "sanitizer $tool rc=$?"; tail -3 gpurun_out/r02a_sanitizer_$tool.log
done
