#!/usr/bin/env bash
# per-source-line profile of the two cold kernels (C2): ncu --set full + source page, joined with the line table of the cubin
set -u
O=gpurun_out; T=${1:-r2t}
mkdir -p $O
NCU="ncu --set full --clock-control none --import-source on"
W=$(python scripts/ncu_target.py C2 1000 run_ahead=0 | awk '{print $3}')
timeout 300 $NCU -k regex:k_wave --launch-skip $((4 * W)) -c 2 -o $O/${T}_c2_cold -f python scripts/ncu_target.py C2 1000 run_ahead=0 > $O/${T}_c2_cold.log 2>&1
ncu -i $O/${T}_c2_cold.ncu-rep --page source --csv 2>/dev/null | gzip > $O/${T}_source_c2_cold.csv.gz
python scripts/ncu_summary.py $O/${T}_c2_cold.ncu-rep $O/${T}_ncu_full_c2_cold.md > /dev/null 2>&1
rm -f $O/${T}_c2_cold.ncu-rep
ls -la $O | grep ${T}
