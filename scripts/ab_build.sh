#!/usr/bin/env bash
# A/B builds of the engine library: .ab/libtrekis3_gpu_<name>.so = the default build plus the given -D switches
# (any macro the sources test; the out-of-line copies of m_div / interp5t / find_lut tried with it in round 2 were all slower,
# profiles/r2k_sweep_outline_builds.txt, and are gone again).  Compare on the GPU box with
#   TRK3_GPU_LIB=.ab/libtrekis3_gpu_<name>.so python scripts/sweep.py C2 1000 ""
# Usage: scripts/ab_build.sh name "-DSOME_SWITCH" [name2 "flags2" ...]
set -eu
cd "$(dirname "$0")/../trekis-3_b200/csrc"
mkdir -p ../../.ab
BASE="-O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -Xcompiler -fPIC -DTRK_MATH_OUTLINE"   # + the switches under test (the Makefile adds -DTRK_PHILOX_UNROLL=5)
[ -f cuda/tables_gpu.o ] || nvcc $BASE -fmad=false -c -o cuda/tables_gpu.o cuda/tables_gpu.cu
while [ $# -ge 2 ]; do
    name=$1; flags=$2; shift 2
    nvcc $BASE $flags -shared -o ../../.ab/libtrekis3_gpu_$name.so cuda/engine.cu common/trk3_layout.c cuda/tables_gpu.o -ldl &
done
wait
ls -la ../../.ab
