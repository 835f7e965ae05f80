#!/usr/bin/env bash
# final build on 2 GPUs: reduced tallies of 2 ranks == the one-GPU run; C2 weak bench line (e2e path collective, staged re-binding)
set -u
O=gpurun_out; T=${1:-r2E}
mkdir -p $O
timeout 200 python -m pytest tests/test_gpu_multi.py -m gpu -q > $O/${T}_pytest_multi2.log 2>&1; tail -2 $O/${T}_pytest_multi2.log
timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29544 bench.py --gpus 2 --steps 5 --warmup 3 --no-cpu-baseline --no-throughput > $O/${T}_bench_C2_2gpu.json 2> $O/${T}_bench_C2_2gpu.err
python -c "
import json; d=json.load(open('$O/${T}_bench_C2_2gpu.json')); print('C2 2gpu', round(d['value']), round(d['ms_per_step'],2), d['scaling'], d['n_gpus'], 'e2e', round(d['e2e']['value']), d['e2e']['ms_per_call'])"
