#!/bin/bash
# One gpurun call, 1 GPU (first call of round 2): the leaf kernel's TMA tile fill (-DNBODY_LEAF_BULK=1), which was written and
# model-checked (tests/test_leaf_fill_model.py) in round 1 after the GPU budget was spent and has NOT run on hardware.
# Build both libraries in the authoring container first (the .so files travel with the snapshot):
#     python nbody_b200/build.py
#     NBODY_BUILD_TAG=bulk NBODY_BUILD_DEFS="-DNBODY_LEAF_BULK=1" python nbody_b200/build.py
# Order: parity of the experimental library under a short timeout (an mbarrier mistake hangs the kernel: the timeout, not gpurun's
# limit, must end it), then A/B bench lines, then one ncu capture of k_leaf from the faster of the two.
mkdir -p gpurun_out
BULK=$PWD/nbody_b200/libnbody_cuda_bulk.so
if [ ! -f "$BULK" ]; then echo "no $BULK: build it before the call"; exit 1; fi
NBODY_CUDA_LIB=$BULK timeout 300 python -m pytest tests/test_golden_fmm.py tests/test_gpu_parity.py -q -m gpu -x > gpurun_out/r02a_parity_bulk.log 2>&1
rc=$?; echo "rc=$rc" >> gpurun_out/r02a_parity_bulk.log; tail -6 gpurun_out/r02a_parity_bulk.log
timeout 150 python bench.py --no-cpu-baseline --no-reference-capacity > gpurun_out/r02a_bench_default.json 2> gpurun_out/r02a_bench_default.err; echo "bench default rc=$?"
if [ $rc -eq 0 ]; then
	NBODY_CUDA_LIB=$BULK timeout 150 python bench.py --no-cpu-baseline --no-reference-capacity > gpurun_out/r02a_bench_bulk.json 2> gpurun_out/r02a_bench_bulk.err; echo "bench bulk rc=$?"
	NBODY_CUDA_LIB=$BULK timeout 400 ncu --set full --clock-control none --import-source on -k regex:k_leaf -c 1 -o gpurun_out/r02a_leaf_bulk \
		python tools/prof_step.py 16777216 1 4 48 > gpurun_out/r02a_ncu_bulk.log 2>&1; echo "ncu rc=$?"
fi
python - <<'PY'
import json
for f in ("r02a_bench_default.json", "r02a_bench_bulk.json"):
    try:
        d = json.load(open("gpurun_out/" + f)); print(f, round(d["ms_per_step"], 3), {k: round(v, 2) for k, v in d["stage_ms"].items()}, round(d["roofline"]["frac"], 4))
    except Exception as e:
        print(f, "unreadable", e)
PY
