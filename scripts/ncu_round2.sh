#!/usr/bin/env bash
# ncu captures of round 2 (run on the GPU box through gpurun; outputs under gpurun_out/, summaries to be copied to profiles/).
#   1. launch list of the bench command (shares of the step; cold-cache, serialised)
#   2. `ncu --set full` of EVERY launch of one batch for C2 (the bench workload), C4 (100 iterations) and C5 (4096 in flight):
#      raw page -> scripts/ncu_to_json.py -> per-kernel-class DRAM traffic, warp instructions per collision, L2 hit rate,
#      FP64 pipe and issue-slot utilisation for exactly the launches whose algorithmic bytes bench.py counts
#   3. per-instruction source page of the C2 capture (hot electron kernel) for scripts/sass_lines.py
# run_ahead=0 makes the number of launches per batch exact (no speculative empty generations).
set -u
OUT=gpurun_out
TAG=${1:-r2h}
NCU="ncu --set full --clock-control none --import-source on"
mkdir -p $OUT
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $OUT/${TAG}_launches_ncu.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > $OUT/${TAG}_ncu_bench.log 2>&1
for spec in "c2 C2 1000" "c4 C4 100" "c5 C1 4096 batch=4096"; do
    set -- $spec; name=$1; shift
    NCU_STATS_OUT=$OUT/${TAG}_stats_$name.json $NCU -o $OUT/${TAG}_$name -f python scripts/ncu_target.py "$@" run_ahead=0 > $OUT/${TAG}_$name.log 2>&1
    ncu -i $OUT/${TAG}_$name.ncu-rep --page raw --csv > $OUT/${TAG}_raw_$name.csv 2>/dev/null
    python scripts/ncu_to_json.py $OUT/${TAG}_raw_$name.csv $OUT/${TAG}_stats_$name.json $OUT/${TAG}_ncu_classes_$name.json >> $OUT/${TAG}_$name.log 2>&1
    python scripts/ncu_summary.py $OUT/${TAG}_$name.ncu-rep $OUT/${TAG}_ncu_full_$name.md > /dev/null 2>&1
    if [ $name = c2 ]; then ncu -i $OUT/${TAG}_$name.ncu-rep --page source --csv 2>/dev/null | gzip > $OUT/${TAG}_source_c2.csv.gz; fi
    rm -f $OUT/${TAG}_$name.ncu-rep
    gzip -f $OUT/${TAG}_raw_$name.csv
done
ls -la $OUT | grep ${TAG}; du -sh $OUT
