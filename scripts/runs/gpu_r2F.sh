#!/usr/bin/env bash
# final build: ncu launch list of the bench command; bench lines of C1 and C3 (100-iteration steps)
set -u
O=gpurun_out; T=${1:-r2F}
mkdir -p $O
timeout 110 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file $O/${T}_launches_ncu.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-throughput > $O/${T}_ncu_bench.log 2>&1; wc -l $O/${T}_launches_ncu.csv
for c in C1 C3; do timeout 60 python bench.py --config $c --steps 5 --warmup 3 --no-cpu-baseline --no-throughput > $O/${T}_bench_$c.json 2> $O/${T}_bench_$c.err; python -c "
import json; d=json.load(open('$O/${T}_bench_$c.json')); print('$c', round(d['value']), round(d['ms_per_step'],2), 'e2e', round(d['e2e']['value']), d['e2e']['ms_per_call'])"; done
