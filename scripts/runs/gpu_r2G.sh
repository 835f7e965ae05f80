#!/usr/bin/env bash
# diagnostic on 2 GPUs: what the collective at the end of a step costs (TRK3_NCCL_PROBE) under NCCL's defaults and two settings
set -u
O=gpurun_out; T=${1:-r2G}
mkdir -p $O
run() { name=$1; shift
  env TRK3_NCCL_PROBE=1 "$@" timeout 50 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29544 bench.py --gpus 2 --steps 5 --warmup 3 --no-cpu-baseline --no-throughput > $O/${T}_bench_$name.json 2> $O/${T}_bench_$name.err
  echo "== $name"; grep "nccl probe" $O/${T}_bench_$name.err | sed -n '7,14p' | cut -c1-250
  python -c "
import json; d=json.load(open('$O/${T}_bench_$name.json')); print('$name', round(d['value']), round(d['ms_per_step'],2), 'e2e', round(d['e2e']['value']))"; }
run default
run ll NCCL_PROTO=LL NCCL_ALGO=Ring
run ch1 NCCL_MAX_NCHANNELS=1 NCCL_MIN_NCHANNELS=1
