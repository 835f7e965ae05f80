#!/usr/bin/env bash
# ncu captures of round 2 (run on the GPU box through gpurun; outputs under gpurun_out/).  run_ahead=0 makes the number of
# launches per batch exact (no speculative empty generations), so that --launch-skip addresses the cold launches.
set -u
OUT=gpurun_out
NCU="ncu --set full --clock-control none --import-source on"
mkdir -p $OUT
# launch list of the bench command (shares of the step; cold-cache, serialised)
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $OUT/r2e_launches_ncu.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > $OUT/r2e_ncu_bench.log 2>&1
# C2: hot electrons / holes of the first generations, then the two cold launches
$NCU -k regex:k_hot -c 4 -o $OUT/r2e_c2_hot -f python scripts/ncu_target.py C2 1000 run_ahead=0 > $OUT/r2e_c2_hot.log 2>&1
W=$(python scripts/ncu_target.py C2 1000 run_ahead=0 | awk '{print $3}')
$NCU -k regex:k_wave --launch-skip $((4 * W)) -c 2 -o $OUT/r2e_c2_cold -f python scripts/ncu_target.py C2 1000 run_ahead=0 > $OUT/r2e_c2_cold.log 2>&1
$NCU -k regex:"k_shi|k_snapshot|k_ion_emit" -c 4 -o $OUT/r2e_c2_misc -f python scripts/ncu_target.py C2 1000 run_ahead=0 > $OUT/r2e_c2_misc.log 2>&1
# C4 (U in Au, 100 iterations): the warm kernels that dominate its step + hot kernels of generation 1
$NCU -k regex:"k_wave|k_hot" --launch-skip 5 -c 5 -o $OUT/r2e_c4_gen1 -f python scripts/ncu_target.py C4 100 run_ahead=0 > $OUT/r2e_c4_gen1.log 2>&1
# C5 (= C1 at 4096 iterations in flight): the two cold launches
W=$(python scripts/ncu_target.py C1 4096 batch=4096 run_ahead=0 | awk '{print $3}')
$NCU -k regex:k_wave --launch-skip $((3 * W)) -c 2 -o $OUT/r2e_c5_cold -f python scripts/ncu_target.py C1 4096 batch=4096 run_ahead=0 > $OUT/r2e_c5_cold.log 2>&1
$NCU -k regex:k_hot -c 2 -o $OUT/r2e_c5_hot -f python scripts/ncu_target.py C1 4096 batch=4096 run_ahead=0 > $OUT/r2e_c5_hot.log 2>&1
# summaries (and the per-instruction source page of the hot electron kernel) stay, the reports themselves are too large to travel
for r in c2_hot c2_cold c2_misc c4_gen1 c5_cold c5_hot; do python scripts/ncu_summary.py $OUT/r2e_$r.ncu-rep $OUT/r2e_ncu_full_$r.md > /dev/null 2>&1; ncu -i $OUT/r2e_$r.ncu-rep --page raw --csv > $OUT/r2e_raw_$r.csv 2>/dev/null; done
ncu -i $OUT/r2e_c2_hot.ncu-rep --page source --csv > $OUT/r2e_source_c2_hot.csv 2>/dev/null
gzip -f $OUT/r2e_source_c2_hot.csv
rm -f $OUT/*.ncu-rep
ls -la $OUT | grep r2e
