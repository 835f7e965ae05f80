#!/usr/bin/env bash
# GPU call: parity suite, ion-slice sweep (option shi_slice) on C2 / C1 / C5-sized batches, trace of one sliced step, bench
set -u
O=gpurun_out; T=${1:-r2l}
mkdir -p $O
timeout 700 python -m pytest tests -m gpu -x -q > $O/${T}_pytest.log 2>&1; tail -3 $O/${T}_pytest.log
timeout 200 python scripts/sweep.py C2 1000 "" "shi_slice=32" "shi_slice=64" "shi_slice=96" "shi_slice=128" "" "shi_slice=48" > $O/${T}_sweep_slice_C2.txt 2>&1; grep min $O/${T}_sweep_slice_C2.txt
timeout 100 python scripts/sweep.py C1 100 "" "shi_slice=16" "shi_slice=32" "shi_slice=64" > $O/${T}_sweep_slice_C1.txt 2>&1; grep min $O/${T}_sweep_slice_C1.txt
timeout 100 python scripts/sweep.py C3 100 "" "shi_slice=16" "shi_slice=32" "shi_slice=64" > $O/${T}_sweep_slice_C3.txt 2>&1; grep min $O/${T}_sweep_slice_C3.txt
timeout 150 python scripts/sweep.py C1 4096 "batch=4096" "batch=4096,shi_slice=32" "batch=4096,shi_slice=64" > $O/${T}_sweep_slice_C5.txt 2>&1; grep min $O/${T}_sweep_slice_C5.txt
timeout 100 python scripts/trace.py C2 1000 shi_slice=64 > $O/${T}_trace_slice64.txt 2>&1; head -30 $O/${T}_trace_slice64.txt | cut -c1-110
timeout 300 python bench.py --steps 5 --warmup 3 > $O/${T}_bench.json 2> $O/${T}_bench.err; head -c 300 $O/${T}_bench.json; echo
