#!/usr/bin/env bash
# A/B builds of the engine library: .ab/libtrekis3_gpu_<name>.so = the default build plus the given -D switches
# (physics.cuh: TRK_OUT_DIV / TRK_OUT_INTERP / TRK_OUT_FIND keep ONE out-of-line copy of a helper in each kernel instead of one
# per call site -- the wave kernels are bound by instruction fetch).  Compare on the GPU box with
#   TRK3_GPU_LIB=.ab/libtrekis3_gpu_<name>.so python scripts/sweep.py C2 1000 ""
# Usage: scripts/ab_build.sh name "-DTRK_OUT_DIV -DTRK_OUT_INTERP" [name2 "flags2" ...]
set -eu
cd "$(dirname "$0")/../trekis-3_b200/csrc"
mkdir -p ../../.ab
BASE="-O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -Xcompiler -fPIC -DTRK_MATH_OUTLINE -DTRK_PHILOX_ROLLED"
[ -f cuda/tables_gpu.o ] || nvcc $BASE -fmad=false -c -o cuda/tables_gpu.o cuda/tables_gpu.cu
while [ $# -ge 2 ]; do
    name=$1; flags=$2; shift 2
    nvcc $BASE $flags -shared -o ../../.ab/libtrekis3_gpu_$name.so cuda/engine.cu common/trk3_layout.c cuda/tables_gpu.o -ldl &
done
wait
ls -la ../../.ab
