#!/usr/bin/env bash
set -u
O=gpurun_out; T=${1:-r2r}
mkdir -p $O
timeout 900 python -m pytest tests -m gpu -q > $O/${T}_pytest.log 2>&1; tail -5 $O/${T}_pytest.log
timeout 300 python bench.py --steps 5 --warmup 3 > $O/${T}_bench.json 2> $O/${T}_bench.err; head -c 300 $O/${T}_bench.json; echo
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > $O/${T}_smoke.log 2>&1; tail -2 $O/${T}_smoke.log
