#!/bin/bash
# Build the product library and the experimental leaf-kernel variants next to it (nbody_b200/libnbody_cuda_<tag>.so, objects under
# nbody_b200/build_<tag>/; all git-ignored, all travel to the GPU box). Load a variant with NBODY_CUDA_LIB=<path>.
set -e
cd "$(dirname "$0")/.."
python nbody_b200/build.py
build() { NBODY_BUILD_TAG=$1 NBODY_BUILD_DEFS="$2" python nbody_b200/build.py > /dev/null; grep -A3 "k_leafILi4ELb1" nbody_b200/build_$1/leaf.o.log | grep -E "Used|spill" | tr '\n' ' '; echo " <- $1"; }
build rowfill "-DNBODY_LEAF_BULK=0"                               # the cp.async row fill of round 1, for A/B runs
build cta3 "-DNBODY_LEAF_MIN_CTAS=3"                              # 3 CTAs of 4 warps per SM at up to 168 registers
build cta3_rows8 "-DNBODY_LEAF_MIN_CTAS=3 -DNBODY_LEAF_ROWS=8"    # ... with 8 source rows in flight per lane
build cta5_g8 "-DNBODY_LEAF_MIN_CTAS=5 -DNBODY_LEAF_G=8"          # 5 CTAs per SM, 8 targets per block
build flat "-DNBODY_LEAF_FLAT=1"                                  # tiles run across the segments of a leaf's source list
