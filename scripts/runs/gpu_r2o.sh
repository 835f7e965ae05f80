#!/usr/bin/env bash
# GPU call: parity suite of the current build (incl. DSF, delta-CDF, BEB), bench
set -u
O=gpurun_out; T=${1:-r2o}
mkdir -p $O
timeout 900 python -m pytest tests -m gpu -q > $O/${T}_pytest.log 2>&1; tail -5 $O/${T}_pytest.log
timeout 300 python bench.py --steps 5 --warmup 3 > $O/${T}_bench.json 2> $O/${T}_bench.err; head -c 300 $O/${T}_bench.json; echo
