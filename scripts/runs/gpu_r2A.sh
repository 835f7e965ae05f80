#!/usr/bin/env bash
# 8-GPU run of the final build: reduced tallies of 2 / 4 / 8 ranks == one-GPU run; C5 strong and C2 weak on 8 GPUs
set -u
O=gpurun_out; T=${1:-r2A}
mkdir -p $O
timeout 400 python -m pytest tests/test_gpu_multi.py -m gpu -q > $O/${T}_pytest_multi8.log 2>&1; tail -2 $O/${T}_pytest_multi8.log
for cfg in "C5 2 1" "C2 5 3"; do set -- $cfg
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29544 bench.py --gpus 8 --config $1 --steps $2 --warmup $3 --no-cpu-baseline > $O/${T}_bench_$1_8gpu.json 2> $O/${T}_bench_$1_8gpu.err
  python -c "
import json; d=json.load(open('$O/${T}_bench_$1_8gpu.json')); print('$1 8gpu', round(d['value']), round(d['ms_per_step'],2), d['scaling'], d['n_gpus'], round(d['e2e']['value']))"
done
