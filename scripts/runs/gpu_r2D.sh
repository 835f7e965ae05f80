#!/usr/bin/env bash
# end-of-round evidence of the final build: parity suite (with the iteration-range edge cases), bench, reference arm, smoke
set -u
O=gpurun_out; T=${1:-r2D}
mkdir -p $O
timeout 300 python bench.py --steps 5 --warmup 3 > $O/${T}_bench.json 2> $O/${T}_bench.err
python -c "
import json; d=json.load(open('$O/${T}_bench.json')); print('value', round(d['value']), 'e2e', round(d['e2e']['value']), d['e2e']['ms_per_call'], d['e2e']['device_ms_per_call'], 'ms', round(d['ms_per_step'],3), 'cpu', d['cpu_baseline']['value'])"
timeout 600 python -m pytest tests -m gpu -q > $O/${T}_pytest.log 2>&1; tail -4 $O/${T}_pytest.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > $O/${T}_smoke.log 2>&1; tail -1 $O/${T}_smoke.log
timeout 200 python bench.py --impl reference --steps 3 --warmup 1 > $O/${T}_bench_ref.json 2> $O/${T}_bench_ref.err; head -c 200 $O/${T}_bench_ref.json; echo
timeout 100 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-throughput > $O/${T}_bench_20steps.json 2> $O/${T}_bench_20steps.err
python -c "
import json; d=json.load(open('$O/${T}_bench_20steps.json')); print('20 steps: value', round(d['value']), 'e2e', round(d['e2e']['value']), d['e2e']['ms_per_call'], d['e2e']['device_ms_per_call'])"
