#!/bin/bash
# One gpurun call, 1 GPU (first call of round 2): the leaf-kernel variants written and model-checked in round 1 after the GPU
# budget was spent — NONE of them has run on hardware:
#   bulk     -DNBODY_LEAF_BULK=1    tile fill with one cp.async.bulk per contiguous source run (tests/test_leaf_fill_model.py)
#   x2       -DNBODY_P2P_F32X2=1    two-wide FP32 interactions (FADD2 / FMUL2 / FFMA2), k_leaf and k_direct
#   bulk_x2  both
#   m2lx2    -DNBODY_M2L_F32X2=1    M2L: the two derivative tensors of the two-interaction (order P-1) path computed two-wide
#   m2lpair  -DNBODY_M2L_PAIR=1     M2L: two sibling targets per warp, everything two-wide (k_m2l_pair)
#   m2lpair2 the same held to 2 CTAs per SM (224 registers, no spill) instead of 3 (168 registers, 20 bytes of spill)
#   all      bulk + x2 + m2lpair
# Build all libraries in the authoring container first (the .so files travel with the snapshot):
#     tools/build_variants.sh
# Order: FFMA2 microbenchmark (does a packed instruction cost one issue slot?), then per variant: parity under a short timeout
# (an mbarrier mistake hangs the kernel: the timeout, not gpurun's limit, must end it), bench line; last, one ncu capture of
# k_leaf from the fastest variant that passed.
mkdir -p gpurun_out
if [ $# -eq 0 ]; then   # the microbenchmark, the one-GPU distributed-sort check and the default bench line only on the first call
if [ -x tools/micro/fma_peak ]; then timeout 120 tools/micro/fma_peak > gpurun_out/r02a_fma_peak.log 2>&1; grep -h "FFMA2\|P2P chain\|dependent" gpurun_out/r02a_fma_peak.log | grep "occ=4\|dependent"; fi
# the distributed sort's device pipeline (slice sorts + merge rounds) on one GPU, default library
NBODY_TEST_EXPERIMENTAL=1 timeout 200 python -m pytest tests/test_gpu_parity.py -q -m gpu -k distributed_sort > gpurun_out/r02a_dist_sort_1gpu.log 2>&1; echo "dist sort 1-GPU rc=$?"; tail -3 gpurun_out/r02a_dist_sort_1gpu.log
NBODY_TEST_EXPERIMENTAL=1 NBODY_VARIANT_LIB=$PWD/nbody_b200/libnbody_cuda_x2.so timeout 100 python -m pytest tests/test_gpu_parity.py -q -m gpu -k two_wide_all_pairs > gpurun_out/r02a_x2_bitwise.log 2>&1; echo "x2 all-pairs bitwise rc=$?"
timeout 150 python bench.py --no-cpu-baseline --no-reference-capacity > gpurun_out/r02a_bench_default.json 2> gpurun_out/r02a_bench_default.err; echo "bench default rc=$?"
fi
TAGS=${@:-bulk bulk_rows2}   # the two-wide variants (x2 bulk_x2 m2lx2 m2lpair m2lpair2 all) only on request: FFMA2 does not pay (profiles/r01o_summary.md); ~2.5 GPU-minutes per tag
for tag in $TAGS; do
	LIB=$PWD/nbody_b200/libnbody_cuda_$tag.so
	if [ ! -f "$LIB" ]; then echo "no $LIB: build it before the call"; continue; fi
	NBODY_CUDA_LIB=$LIB timeout 300 python -m pytest tests/test_golden_fmm.py tests/test_gpu_parity.py -q -m gpu -x > gpurun_out/r02a_parity_$tag.log 2>&1
	rc=$?; echo "rc=$rc" >> gpurun_out/r02a_parity_$tag.log; echo "== $tag parity rc=$rc"; tail -4 gpurun_out/r02a_parity_$tag.log
	if [ $rc -eq 0 ]; then
		NBODY_CUDA_LIB=$LIB timeout 150 python bench.py --no-cpu-baseline --no-reference-capacity > gpurun_out/r02a_bench_$tag.json 2> gpurun_out/r02a_bench_$tag.err; echo "bench $tag rc=$?"
	fi
done
BEST=$(python - <<'PY'
import json
best, best_ms = "default", 1e9
for tag in ("default", "bulk", "bulk_rows2", "x2", "bulk_x2", "m2lx2", "m2lpair", "m2lpair2", "all"):
    try:
        d = json.load(open(f"gpurun_out/r02a_bench_{tag}.json"))
        ms = d["stage_ms"]["ms_leaf"]
        import sys
        print(tag, round(d["ms_per_step"], 3), {k: round(v, 2) for k, v in d["stage_ms"].items()}, "leaf frac", round(d["roofline"]["frac"], 4),
              "all-pairs frac", round(d["p2p_fp32_tflops"]["all_pairs_frac_of_peak"], 4), file=sys.stderr)
        if ms < best_ms and tag in ("default", "bulk", "bulk_rows2", "x2", "bulk_x2"):
            best, best_ms = tag, ms
    except Exception as e:
        import sys
        print(tag, "unreadable", e, file=sys.stderr)
print(best)
PY
)
echo "fastest leaf kernel: $BEST"
[ $# -ne 0 ] && exit 0   # the ncu capture belongs to the first call (leaf variants)
LIB=$PWD/nbody_b200/libnbody_cuda.so; [ "$BEST" != default ] && LIB=$PWD/nbody_b200/libnbody_cuda_$BEST.so
NBODY_CUDA_LIB=$LIB timeout 400 ncu --set full --clock-control none --import-source on -k regex:k_leaf -c 1 -o gpurun_out/r02a_leaf_$BEST \
	python tools/prof_step.py 16777216 1 4 48 > gpurun_out/r02a_ncu_$BEST.log 2>&1; echo "ncu rc=$?"
