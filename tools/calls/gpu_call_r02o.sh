#!/bin/bash
# 1 GPU: what the driver runs at the end of the round — the whole GPU suite, smoke(), the default bench line, the reference arm.
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -x -q -m gpu > gpurun_out/r02o_gpu_suite.log 2>&1; echo "gpu suite rc=$?"; tail -4 gpurun_out/r02o_gpu_suite.log | cut -c1-300
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02o_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/r02o_smoke.log | cut -c1-300
( time timeout 600 python bench.py > gpurun_out/r02o_bench.json 2> gpurun_out/r02o_bench.err ) 2>&1 | grep real; echo "bench rc=$?"
( time timeout 600 python bench.py --impl reference > gpurun_out/r02o_bench_reference.json 2> gpurun_out/r02o_bench_reference.err ) 2>&1 | grep real
python - <<'PY'
import json
d = json.load(open("gpurun_out/r02o_bench.json")); r = json.load(open("gpurun_out/r02o_bench_reference.json"))
print(round(d["ms_per_step"], 3), f'{d["value"]:.4g}', {k: round(v, 2) for k, v in d["stage_ms"].items() if v}, "roofline", round(d["roofline"]["frac"], 4), "accuracy", d["accuracy"]["rms_rel"], "e2e", f'{d["e2e"]["value"]:.4g}', "launches", d["gpu_launches"])
print("reference_capacity", d["reference_capacity"]["ms_per_step"], "config1 ours", f'{d["config1"]["value"]:.4g}', "reference", f'{r["config1"]["value"]:.4g}', "ref value", f'{r["value"]:.4g}', "cpu_baseline", f'{d["cpu_baseline"]["value"]:.4g}')
PY
