#!/usr/bin/env bash
# end-of-round evidence of the final build: parity suite, ncu per-class counters + launch list, bench (which reads the fresh
# profiles/ncu_classes_C2.json), reference arm, smoke, all-configuration times
set -u
O=gpurun_out; T=${1:-r2v}
mkdir -p $O
timeout 900 python -m pytest tests -m gpu -q > $O/${T}_pytest.log 2>&1; tail -4 $O/${T}_pytest.log
bash scripts/ncu_round2.sh $T "list classes" > $O/${T}_ncu.log 2>&1; tail -3 $O/${T}_ncu.log
cp $O/${T}_ncu_classes_C2.json profiles/ncu_classes_C2.json; cp $O/${T}_ncu_classes_C5.json profiles/ncu_classes_C5.json
timeout 300 python bench.py --steps 5 --warmup 3 > $O/${T}_bench.json 2> $O/${T}_bench.err; head -c 300 $O/${T}_bench.json; echo
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > $O/${T}_bench_20steps.json 2> $O/${T}_bench_20steps.err; head -c 200 $O/${T}_bench_20steps.json; echo
timeout 200 python bench.py --impl reference --steps 3 --warmup 1 > $O/${T}_bench_ref.json 2> $O/${T}_bench_ref.err; head -c 200 $O/${T}_bench_ref.json; echo
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > $O/${T}_smoke.log 2>&1; tail -1 $O/${T}_smoke.log
for c in C1 C3 C4 C5; do timeout 400 python bench.py --config $c --steps 3 --warmup 3 --no-cpu-baseline > $O/${T}_bench_$c.json 2> $O/${T}_bench_$c.err; python -c "
import json,sys; d=json.load(open('$O/${T}_bench_$c.json')); print('$c', round(d['value']), round(d['ms_per_step'],2), d['scaling'])"; done
