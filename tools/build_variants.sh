#!/bin/bash
# Build the product library and the experimental leaf-kernel variants next to it (nbody_b200/libnbody_cuda_<tag>.so, objects under
# nbody_b200/build_<tag>/; all git-ignored, all travel to the GPU box). Load a variant with NBODY_CUDA_LIB=<path>.
set -e
cd "$(dirname "$0")/.."
python nbody_b200/build.py
NBODY_BUILD_TAG=bulk NBODY_BUILD_DEFS="-DNBODY_LEAF_BULK=1" python nbody_b200/build.py
NBODY_BUILD_TAG=bulk_rows2 NBODY_BUILD_DEFS="-DNBODY_LEAF_BULK=1 -DNBODY_LEAF_ROWS=2" python nbody_b200/build.py   # tiles padded to 64 instead of 128 sources
(cd tools/micro && nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o fma_peak fma_peak.cu)
for t in bulk bulk_rows2; do grep -A3 "k_leafILi4ELb1" nbody_b200/build_$t/leaf.o.log | grep -E "Used|spill" | tr '\n' ' '; echo " <- $t"; done
