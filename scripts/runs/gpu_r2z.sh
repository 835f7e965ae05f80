#!/usr/bin/env bash
# GPU call: parity suite + step times with one event-counter increment per warp instruction
set -u
O=gpurun_out; T=${1:-r2z}
mkdir -p $O
timeout 900 python -m pytest tests -m gpu -q > $O/${T}_pytest.log 2>&1; tail -5 $O/${T}_pytest.log
timeout 200 python scripts/sweep.py C2 1000 "" "" "" > $O/${T}_sweep_C2.txt 2>&1; grep -A1 min $O/${T}_sweep_C2.txt | cut -c1-330
timeout 100 python scripts/sweep.py C1 100 "" "" > $O/${T}_sweep_C1.txt 2>&1; grep min $O/${T}_sweep_C1.txt
timeout 150 python scripts/sweep.py C1 4096 "batch=4096" "batch=4096" > $O/${T}_sweep_C5.txt 2>&1; grep -A1 min $O/${T}_sweep_C5.txt | cut -c1-330
timeout 150 python scripts/sweep.py C4 100 "" > $O/${T}_sweep_C4.txt 2>&1; grep min $O/${T}_sweep_C4.txt
