#!/usr/bin/env bash
# one GPU call of round 2: parity suite, bench, A/B sweeps (phased cold electrons, out-of-line helper builds), bounded ncu passes
set -u
O=gpurun_out; T=${1:-r2k}
mkdir -p $O
timeout 600 python -m pytest tests -m gpu -x -q > $O/${T}_pytest.log 2>&1; tail -3 $O/${T}_pytest.log
timeout 300 python bench.py --steps 5 --warmup 3 > $O/${T}_bench.json 2> $O/${T}_bench.err; head -c 400 $O/${T}_bench.json; echo
timeout 200 python scripts/sweep.py C2 1000 "" "cold_phased=1" "l2_persist=1" "" "cold_phased=1" "l2_persist=1" "cold_phased=1,l2_persist=1" > $O/${T}_sweep_phased.txt 2>&1; grep min $O/${T}_sweep_phased.txt
timeout 120 python scripts/sweep.py C1 4096 "batch=4096" "batch=4096,cold_phased=1" "batch=4096,l2_persist=1" >> $O/${T}_sweep_phased.txt 2>&1; grep min $O/${T}_sweep_phased.txt | tail -3
for v in div interp find all3; do
    [ -f .ab/libtrekis3_gpu_$v.so ] || continue
    echo "== $v" >> $O/${T}_sweep_ab.txt
    TRK3_GPU_LIB=.ab/libtrekis3_gpu_$v.so timeout 100 python scripts/sweep.py C2 1000 "" >> $O/${T}_sweep_ab.txt 2>&1
done
grep "==\|min" $O/${T}_sweep_ab.txt
bash scripts/ncu_round2.sh $T "list classes" 2>&1 | tail -5
