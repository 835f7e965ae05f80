#!/usr/bin/env bash
# staged table re-binding (one DMA out of a pinned mirror): parity suite, e2e A/B, bench, smoke
set -u
O=gpurun_out; T=${1:-r2B}
mkdir -p $O
timeout 200 python scripts/e2e_ab.py C2 1000 8 > $O/${T}_e2e_ab_C2.txt 2>&1; cat $O/${T}_e2e_ab_C2.txt | tail -5
timeout 100 python scripts/e2e_ab.py C1 100 8 > $O/${T}_e2e_ab_C1.txt 2>&1; tail -4 $O/${T}_e2e_ab_C1.txt
timeout 300 python bench.py --steps 5 --warmup 3 > $O/${T}_bench.json 2> $O/${T}_bench.err; head -c 300 $O/${T}_bench.json; echo
python -c "
import json; d=json.load(open('$O/${T}_bench.json')); print('value', round(d['value']), 'e2e', round(d['e2e']['value']), 'ms', round(d['ms_per_step'],3), 'cpu', d['cpu_baseline']['value'])"
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > $O/${T}_smoke.log 2>&1; tail -1 $O/${T}_smoke.log
timeout 900 python -m pytest tests -m gpu -q -x > $O/${T}_pytest.log 2>&1; tail -4 $O/${T}_pytest.log
