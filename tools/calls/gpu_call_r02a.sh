#!/bin/bash
# One gpurun call, 1 GPU (first call of round 2): what was written and model-checked in round 1 after the GPU budget was spent and
# has NOT run on hardware:
#   bulk        -DNBODY_LEAF_BULK=1                      leaf kernel: tile fill with one cp.async.bulk per contiguous source run
#   bulk_rows2  ... -DNBODY_LEAF_ROWS=2                  the same with 64-source instead of 128-source padding granularity
#   nbody_cuda_sort_runs                                 the device side of the distributed sort (slice sorts + merge rounds) on one GPU
# Build the libraries in the authoring container first (the .so files travel with the snapshot):  tools/build_variants.sh
# Order: distributed-sort pipeline check, default bench line, then per variant: parity under a short timeout (an mbarrier mistake
# hangs the kernel: the timeout, not gpurun's limit, must end it) and a bench line; last, one ncu capture of k_leaf from the fastest.
mkdir -p gpurun_out
NBODY_TEST_EXPERIMENTAL=1 timeout 200 python -m pytest tests/test_gpu_parity.py -q -m gpu -k distributed_sort > gpurun_out/r02a_dist_sort_1gpu.log 2>&1; echo "dist sort 1-GPU rc=$?"; tail -3 gpurun_out/r02a_dist_sort_1gpu.log
timeout 150 python bench.py --no-cpu-baseline --no-reference-capacity > gpurun_out/r02a_bench_default.json 2> gpurun_out/r02a_bench_default.err; echo "bench default rc=$?"
TAGS=${@:-bulk bulk_rows2}
for tag in $TAGS; do
	LIB=$PWD/nbody_b200/libnbody_cuda_$tag.so
	if [ ! -f "$LIB" ]; then echo "no $LIB: build it before the call"; continue; fi
	NBODY_CUDA_LIB=$LIB timeout 300 python -m pytest tests/test_golden_fmm.py tests/test_gpu_parity.py -q -m gpu -x > gpurun_out/r02a_parity_$tag.log 2>&1
	rc=$?; echo "rc=$rc" >> gpurun_out/r02a_parity_$tag.log; echo "== $tag parity rc=$rc"; tail -4 gpurun_out/r02a_parity_$tag.log
	if [ $rc -eq 0 ]; then
		NBODY_CUDA_LIB=$LIB timeout 150 python bench.py --no-cpu-baseline --no-reference-capacity > gpurun_out/r02a_bench_$tag.json 2> gpurun_out/r02a_bench_$tag.err; echo "bench $tag rc=$?"
	fi
done
BEST=$(TAGS="$TAGS" python - <<'PY'
import json, os, sys
best, best_ms = "default", 1e9
for tag in ["default"] + os.environ["TAGS"].split():
    try:
        d = json.load(open(f"gpurun_out/r02a_bench_{tag}.json"))
        ms = d["stage_ms"]["ms_leaf"]
        print(tag, round(d["ms_per_step"], 3), {k: round(v, 2) for k, v in d["stage_ms"].items()}, "leaf frac", round(d["roofline"]["frac"], 4), file=sys.stderr)
        if ms < best_ms:
            best, best_ms = tag, ms
    except Exception as e:
        print(tag, "unreadable", e, file=sys.stderr)
print(best)
PY
)
echo "fastest leaf kernel: $BEST"
LIB=$PWD/nbody_b200/libnbody_cuda.so; [ "$BEST" != default ] && LIB=$PWD/nbody_b200/libnbody_cuda_$BEST.so
NBODY_CUDA_LIB=$LIB timeout 400 ncu --set full --clock-control none --import-source on -k regex:k_leaf -c 1 -o gpurun_out/r02a_leaf_$BEST \
	python tools/prof_step.py 16777216 1 4 48 > gpurun_out/r02a_ncu_$BEST.log 2>&1; echo "ncu rc=$?"
