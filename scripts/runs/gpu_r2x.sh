#!/usr/bin/env bash
set -u
O=gpurun_out; T=${1:-r2x}
mkdir -p $O
timeout 200 python scripts/sweep.py C2 1000 "" "use_smem=0" "" "use_smem=0" > $O/${T}_sweep_smem_C2.txt 2>&1; grep -A1 min $O/${T}_sweep_smem_C2.txt | cut -c1-320
timeout 150 python scripts/sweep.py C1 4096 "batch=4096" "batch=4096,use_smem=0" > $O/${T}_sweep_smem_C5.txt 2>&1; grep -A1 min $O/${T}_sweep_smem_C5.txt | cut -c1-320
