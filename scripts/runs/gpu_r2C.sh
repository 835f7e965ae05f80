#!/usr/bin/env bash
# staged table re-binding, second run (handle destruction fixed, pinned mirror allocated with the handle): parity suite, e2e A/B, bench
set -u
O=gpurun_out; T=${1:-r2C}
mkdir -p $O
timeout 200 python scripts/e2e_ab.py C2 1000 8 > $O/${T}_e2e_ab_C2.txt 2>&1; tail -5 $O/${T}_e2e_ab_C2.txt
timeout 600 python -m pytest tests -m gpu -q > $O/${T}_pytest.log 2>&1; tail -4 $O/${T}_pytest.log
timeout 300 python bench.py --steps 5 --warmup 3 > $O/${T}_bench.json 2> $O/${T}_bench.err
python -c "
import json; d=json.load(open('$O/${T}_bench.json')); print('value', round(d['value']), 'e2e', round(d['e2e']['value']), 'ms', round(d['ms_per_step'],3), 'cpu', d['cpu_baseline']['value'])"
