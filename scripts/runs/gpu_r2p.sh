#!/usr/bin/env bash
# GPU call: parity suite with the barrier-free closing of the hot kernels, A/B of it, ncu of the hot launches, bench
set -u
O=gpurun_out; T=${1:-r2p}
mkdir -p $O
timeout 900 python -m pytest tests -m gpu -q > $O/${T}_pytest.log 2>&1; tail -5 $O/${T}_pytest.log
timeout 200 python scripts/sweep.py C2 1000 "" "" "" > $O/${T}_sweep_C2.txt 2>&1; grep min $O/${T}_sweep_C2.txt
timeout 100 python scripts/sweep.py C1 100 "" "" > $O/${T}_sweep_C1.txt 2>&1; grep min $O/${T}_sweep_C1.txt
timeout 150 python scripts/sweep.py C1 4096 "batch=4096" "batch=4096" > $O/${T}_sweep_C5.txt 2>&1; grep min $O/${T}_sweep_C5.txt
NCU="ncu --set full --clock-control none --import-source on"
timeout 240 $NCU -k regex:k_hot -c 3 -o $O/${T}_c2_hot -f python scripts/ncu_target.py C2 1000 run_ahead=0 > $O/${T}_c2_hot.log 2>&1
python scripts/ncu_summary.py $O/${T}_c2_hot.ncu-rep $O/${T}_ncu_full_c2_hot.md > /dev/null 2>&1; rm -f $O/${T}_c2_hot.ncu-rep
grep -E "^## |duration|barrier" $O/${T}_ncu_full_c2_hot.md | head -12
timeout 300 python bench.py --steps 5 --warmup 3 > $O/${T}_bench.json 2> $O/${T}_bench.err; head -c 300 $O/${T}_bench.json; echo
