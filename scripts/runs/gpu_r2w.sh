#!/usr/bin/env bash
# 2-GPU check of the final build: reduced tallies == one-GPU run (NCCL inside the library), C5 strong and C2 weak on 1 and 2 GPUs
set -u
O=gpurun_out; T=${1:-r2w}
mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_multi.py -m gpu -q > $O/${T}_pytest_multi2.log 2>&1; tail -2 $O/${T}_pytest_multi2.log
timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > $O/${T}_bench_c2_1gpu.json 2> $O/${T}_bench_c2_1gpu.err
python -c "
import json; d=json.load(open('$O/${T}_bench_c2_1gpu.json')); print('c2 1gpu', round(d['value']), d['throughput'])"
timeout 300 python bench.py --config C5 --steps 2 --warmup 1 --no-cpu-baseline > $O/${T}_bench_c5_strong_1gpu.json 2> $O/${T}_bench_c5_strong_1gpu.err
for cfg in "C5 2 1" "C2 5 3"; do set -- $cfg
  timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --config $1 --steps $2 --warmup $3 --no-cpu-baseline > $O/${T}_bench_$1_2gpu.json 2> $O/${T}_bench_$1_2gpu.err
done
for f in c5_strong_1gpu C5_2gpu C2_2gpu; do python -c "
import json; d=json.load(open('$O/${T}_bench_$f.json')); print('$f', round(d['value']), round(d['ms_per_step'],2), d['scaling'], d['n_gpus'], round(d['e2e']['value']))"; done
