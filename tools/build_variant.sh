#!/bin/bash
# Build a measurement variant of the library next to the product one (never loaded by the product path):
#   tools/build_variant.sh NAME "-DPB2_HINGE_KO=7"   ->  tools/ab/lib_NAME.so   (git-ignored, travels with gpurun)
# Compile-time hooks in csrc/sim.cu: PB2_HINGE_KO (1 no column counts, 2 no loss sum, 4 no row counts: wrong results,
# timing only), PB2_HINGE_PIPES (1 column indicator on the ALU pipe, 2 / 4 row / rank indicator on the FMA pipe:
# identical results), PB2_STAGE_CAP (cap on the TMA pipeline stages).  Time with tools/ab_hinge.py inside ONE gpurun
# call, alternating with the product library.
set -e
ROOT=$(cd "$(dirname "$0")/.." && pwd)
python -m peppa_b200.build > /dev/null
mkdir -p "$ROOT/tools/ab"
cd "$ROOT/peppa_b200/csrc"
nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -I ../../include -I . -DPB2_MEASURE $2 \
    -c sim.cu -o "build_measure/sim_$1.variant.o"
cd build_measure      # variants are measurement builds: they link the -DPB2_MEASURE objects (pb2_debug_* selectors)
nvcc -shared -gencode arch=compute_100a,code=sm_100a -o "$ROOT/tools/ab/lib_$1.so" \
    host_util.o triplet.o rowstats.o "sim_$1.variant.o" gradgemm.o step.o proj.o collective.o sampler.o -ldl
rm -f "sim_$1.variant.o"
echo "$ROOT/tools/ab/lib_$1.so"
