#!/usr/bin/env bash
# ncu captures of round 2 (run on the GPU box through gpurun; outputs under gpurun_out/, summaries to be copied to profiles/).
#   1. launch list of the bench command (shares of the step; cold-cache, serialised)
#   2. EVERY launch of one batch of C2 (the bench workload) with the dozen counters bench.py's roofline needs -- DRAM bytes, warp
#      instructions, L2/L1 hit rates, FP64 pipe, issue slots, lanes per instruction, the no_instruction and barrier stalls:
#      raw page -> scripts/ncu_to_json.py -> per-kernel-class figures for exactly the launches whose algorithmic bytes bench.py
#      counts (profiles/ncu_classes_C2.json).  A metric list instead of `--set full`: ~5 replay passes per launch instead of ~40
#      (the first version of this script ran `--set full` over every launch of C2, C4 and C5 and did not finish in 40 minutes).
#   3. `--set full` of a FEW launches per question: C2 hot generations 0-1, C2 cold launches,
#      C4 generation 1, C5 cold launches; summaries by scripts/ncu_summary.py, per-instruction source page of the C2 hot kernel.
# run_ahead=0 makes the number of launches per batch exact (no speculative empty generations), so that --launch-skip addresses
# the cold launches.  Every stage runs under its own `timeout`.
set -u
OUT=gpurun_out
TAG=${1:-r2k}
STAGES=${2:-"list classes full"}
NCU="ncu --set full --clock-control none --import-source on"
METRICS=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__inst_executed.sum,lts__t_sector_hit_rate.pct,l1tex__t_sector_hit_rate.pct,sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active,smsp__issue_active.avg.pct_of_peak_sustained_active,smsp__thread_inst_executed_per_inst_executed.ratio,smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio,smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio
mkdir -p $OUT
for stage in $STAGES; do case $stage in
list)
    timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $OUT/${TAG}_launches_ncu.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-throughput > $OUT/${TAG}_ncu_bench.log 2>&1
    ;;
classes)
    for spec in "C2 C2 1000" "C5 C1 4096 batch=4096"; do
        set -- $spec; name=$1; shift
        NCU_STATS_OUT=$OUT/${TAG}_stats_$name.json timeout 420 ncu --metrics $METRICS --clock-control none --csv --page raw --log-file $OUT/${TAG}_raw_$name.csv \
            python scripts/ncu_target.py "$@" run_ahead=0 > $OUT/${TAG}_classes_$name.log 2>&1
        grep -v "^==" $OUT/${TAG}_raw_$name.csv > $OUT/${TAG}_raw_$name.clean.csv
        python scripts/ncu_to_json.py $OUT/${TAG}_raw_$name.clean.csv $OUT/${TAG}_stats_$name.json $OUT/${TAG}_ncu_classes_$name.json >> $OUT/${TAG}_classes_$name.log 2>&1
        rm -f $OUT/${TAG}_raw_$name.clean.csv; gzip -f $OUT/${TAG}_raw_$name.csv
    done
    ;;
full)
    T="timeout 240"
    $T $NCU -k regex:k_hot -c 3 -o $OUT/${TAG}_c2_hot -f python scripts/ncu_target.py C2 1000 run_ahead=0 > $OUT/${TAG}_c2_hot.log 2>&1
    W=$(python scripts/ncu_target.py C2 1000 run_ahead=0 | awk '{print $3}')
    $T $NCU -k regex:k_wave --launch-skip $((4 * W)) -c 2 -o $OUT/${TAG}_c2_cold -f python scripts/ncu_target.py C2 1000 run_ahead=0 > $OUT/${TAG}_c2_cold.log 2>&1
    $T $NCU -k regex:"k_wave|k_hot" --launch-skip 5 -c 5 -o $OUT/${TAG}_c4_gen1 -f python scripts/ncu_target.py C4 100 run_ahead=0 > $OUT/${TAG}_c4_gen1.log 2>&1
    for r in c2_hot c2_cold c4_gen1; do
        [ -f $OUT/${TAG}_$r.ncu-rep ] || continue
        python scripts/ncu_summary.py $OUT/${TAG}_$r.ncu-rep $OUT/${TAG}_ncu_full_$r.md > /dev/null 2>&1
    done
    [ -f $OUT/${TAG}_c2_hot.ncu-rep ] && ncu -i $OUT/${TAG}_c2_hot.ncu-rep --page source --csv 2>/dev/null | gzip > $OUT/${TAG}_source_c2_hot.csv.gz
    rm -f $OUT/${TAG}_*.ncu-rep
    ;;
esac; done
ls -la $OUT | grep ${TAG}; du -sh $OUT
