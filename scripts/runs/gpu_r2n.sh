#!/usr/bin/env bash
# GPU call: the one GPU test that changed, C4 (U in Au) option sweep: warm slice / warm threshold / hot slices
set -u
O=gpurun_out; T=${1:-r2n}
mkdir -p $O
timeout 300 python -m pytest tests/test_gpu.py -m gpu -q -k "beb" > $O/${T}_pytest_beb.log 2>&1; tail -3 $O/${T}_pytest_beb.log
timeout 400 python scripts/sweep.py C4 100 "" "warm_slice=128" "warm_slice=256" "warm_slice=512" "warm_pinel=0.3" "warm_pinel=0.7" "warm_pinel=0.9" "warm_slice=256,warm_pinel=0.7" "hot_slice=128" "warm_holes=0" > $O/${T}_sweep_C4.txt 2>&1; grep min $O/${T}_sweep_C4.txt
timeout 100 python scripts/trace.py C4 100 > $O/${T}_trace_C4.txt 2>&1; grep -c trace $O/${T}_trace_C4.txt
