#!/bin/bash
# 1 GPU: block-granular leaf tickets (parity, bench), whole GPU suite, config 5's single-rank point, sanitizer on the partitioned path,
# ncu launch list of one step.
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/r02j_gpu_suite.log 2>&1; echo "gpu suite rc=$?"; tail -5 gpurun_out/r02j_gpu_suite.log | cut -c1-300
timeout 400 python bench.py --no-reference-capacity > gpurun_out/r02j_bench.json 2> gpurun_out/r02j_bench.err; echo "bench rc=$?"; tail -2 gpurun_out/r02j_bench.err
python - <<'PY'
import json
d = json.load(open("gpurun_out/r02j_bench.json")); print(round(d["ms_per_step"], 3), {k: round(v, 2) for k, v in d["stage_ms"].items() if v}, "leaf frac", round(d["p2p_fp32_tflops"]["tree_p2p_frac_of_peak"], 4), d["accuracy"]["rms_rel"], d["clocks"])
PY
timeout 600 python bench.py --workload uniform --particles 134217728 --leaf-capacity 80 --steps 3 --warmup 3 --no-cpu-baseline --no-reference-capacity --no-config1 --accuracy-targets 8192 > gpurun_out/r02j_config5_uniform_2p27_1gpu.json 2> gpurun_out/r02j_config5_1gpu.err; echo "config5 1gpu rc=$?"; tail -2 gpurun_out/r02j_config5_1gpu.err
python - <<'PY'
import json
d = json.load(open("gpurun_out/r02j_config5_uniform_2p27_1gpu.json")); print("config5 1 rank", round(d["ms_per_step"], 2), {k: round(v, 2) for k, v in d["stage_ms"].items() if v}, d["accuracy"], d["counts"]["device_bytes"])
PY
cat > /tmp/virt.py <<'PY'
import sys; sys.path.insert(0, ".")
import nbody_b200
from nbody_b200 import workloads
P = workloads.plummer(20000)
g = nbody_b200.VirtualGroup([1, 1, 1], P, 1e-3, 4, leaf_capacity=8)
for _ in range(2): g.step()
print([ (s["n_particles"], s["halo_particles"], s["imported_nodes"], s["migrated_particles"]) for s in g.stats()])
g.close()
PY
for tool in memcheck racecheck; do
	timeout 600 compute-sanitizer --tool $tool --error-exitcode 9 python /tmp/virt.py > gpurun_out/r02j_sanitizer_partitioned_$tool.log 2>&1; echo "sanitizer $tool rc=$?"; tail -3 gpurun_out/r02j_sanitizer_partitioned_$tool.log | cut -c1-300
done
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02j_launches.csv python tools/prof_step.py 16777216 2 4 48 > gpurun_out/r02j_launches.log 2>&1; echo "ncu launch list rc=$?"
