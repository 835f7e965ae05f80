#!/usr/bin/env bash
# GPU call: parity suite of the current build, bench (with profiles/ncu_classes_C2.json present), --set full of a few launches
set -u
O=gpurun_out; T=${1:-r2m}
mkdir -p $O
timeout 900 python -m pytest tests -m gpu -q > $O/${T}_pytest.log 2>&1; tail -5 $O/${T}_pytest.log
timeout 300 python bench.py --steps 5 --warmup 3 > $O/${T}_bench.json 2> $O/${T}_bench.err; head -c 300 $O/${T}_bench.json; echo
timeout 200 python bench.py --impl reference --steps 3 --warmup 1 > $O/${T}_bench_ref.json 2> $O/${T}_bench_ref.err; head -c 300 $O/${T}_bench_ref.json; echo
bash scripts/ncu_round2.sh $T "full" 2>&1 | tail -12
