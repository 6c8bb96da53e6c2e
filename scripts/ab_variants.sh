#!/bin/bash
# Scratch: build variant libraries for an A/B on the GPU box (they travel with the gpurun snapshot; *.so is git-ignored).
#   scripts/ab_variants.sh unroll2:-DMCDP_QUAD_UNIT_UNROLL=2 w18:-DMCDP_QUAD_MAX_THREADS=576
# then on the box:  for f in scratch_libs/*.so; do MCDP_LIB=$f python scripts/ab_quad3.py; done
set -e
cd "$(dirname "$0")/.."
mkdir -p scratch_libs
cp mc_dagprop_b200/libmcdp_b200.so scratch_libs/libbase.so
for v in "$@"; do
  name=${v%%:*}; flags=${v#*:}
  nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -ccbin ${MCDP_HOST_CXX:-/usr/bin/g++} \
       -Xcompiler -fPIC,-O3 -shared -cudart static ${flags//,/ } -o scratch_libs/lib$name.so \
       mc_dagprop_b200/csrc/mcdp_capi.cu mc_dagprop_b200/csrc/mcdp_analytic.cu mc_dagprop_b200/csrc/mcdp_plan.cpp &
done
wait
ls -la scratch_libs
