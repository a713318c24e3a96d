#!/bin/bash
# Two-GPU check (run with: /usr/local/graft/bin/gpurun --gpus 2 --timeout 900 -- 'bash scripts/gpu_checkup_2gpu.sh'):
# the default data-parallel step (one all-reduce after backward) against the gradient buckets exchanged during backward, with and
# without the side streams of the coarse pyramid levels.
set -u
mkdir -p gpurun_out
run() { timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $((29500 + RANDOM % 400)) \
          bench.py --gpus 2 --steps 10 --warmup 3 --no-cpu-baseline "$@"; }
for e in "" buckets no_streams no_streams,buckets ""; do
  n=$(echo "$e" | tr ',' '_'); [ -z "$n" ] && n=default
  run --experimental "$e" > gpurun_out/r2g2_$n.json 2>> gpurun_out/r2g2.err
  python -c "import json,sys; d=json.load(open('gpurun_out/r2g2_$n.json')); print('$n', d['ms_per_step'], d['value'], d['e2e']['ms_per_step'])"
done
