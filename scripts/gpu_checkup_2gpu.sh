#!/bin/bash
# Two-GPU check (run with: /usr/local/graft/bin/gpurun --gpus 2 --timeout 900 -- 'bash scripts/gpu_checkup_2gpu.sh'):
# the default data-parallel step (one all-reduce after backward) against the opt-in gradient buckets exchanged during backward.
set -u
mkdir -p gpurun_out
run() { timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29531 \
          bench.py --gpus 2 --steps 10 --warmup 3 "$@"; }
run > gpurun_out/bench_2gpu.json 2> gpurun_out/bench_2gpu.err
run --experimental buckets > gpurun_out/bench_2gpu_buckets.json 2>> gpurun_out/bench_2gpu.err
head -c 600 gpurun_out/bench_2gpu.json; echo; head -c 600 gpurun_out/bench_2gpu_buckets.json; echo
