#!/usr/bin/env bash
# A/B of the Philox builds (rounds per loop trip, out-of-line copy) on C2 and on a 4096-iteration batch (C5); phi-wrap change is in all
set -u
O=gpurun_out; T=${1:-r2u}
mkdir -p $O
echo "== default (rolled, mul.wide)" > $O/${T}_sweep_philox.txt
timeout 100 python scripts/sweep.py C2 1000 "" >> $O/${T}_sweep_philox.txt 2>&1
timeout 100 python scripts/sweep.py C1 4096 "batch=4096" >> $O/${T}_sweep_philox.txt 2>&1
for v in u1 u2 u5 u10 out10 out2; do
    echo "== $v" >> $O/${T}_sweep_philox.txt
    TRK3_GPU_LIB=.ab/libtrekis3_gpu_$v.so timeout 100 python scripts/sweep.py C2 1000 "" >> $O/${T}_sweep_philox.txt 2>&1
    TRK3_GPU_LIB=.ab/libtrekis3_gpu_$v.so timeout 100 python scripts/sweep.py C1 4096 "batch=4096" >> $O/${T}_sweep_philox.txt 2>&1
done
grep "==\|min" $O/${T}_sweep_philox.txt | cut -c1-130
