#!/bin/bash
# One 8-GPU box (gpurun --gpus 8): gradient-bucket overlap A/B over NCCL, BASELINE config 3 (LIDC-shaped,
# 8 x B200 data parallel) and config 4 (sliding-window inference, tiles sharded over 2 / 4 / 8 ranks).  Outputs: gpurun_out/r2m8_*.json
set -u
mkdir -p gpurun_out
tr() { local n=$1; shift; timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29500 + RANDOM % 400)) "$@"; }
tr 8 bench.py --gpus 8 --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2m8_bench8.json 2> gpurun_out/r2m8.err
tr 8 bench.py --gpus 8 --steps 20 --warmup 5 --no-cpu-baseline --experimental buckets > gpurun_out/r2m8_bench8_buckets.json 2>> gpurun_out/r2m8.err
tr 8 bench.py --gpus 8 --steps 8 --warmup 3 --no-cpu-baseline --config lidc > gpurun_out/r2m8_lidc8.json 2>> gpurun_out/r2m8.err
for n in 2 4 8; do
  tr $n scripts/bench_inference.py --reps 2 > gpurun_out/r2m8_infer$n.json 2>> gpurun_out/r2m8.err
done
for f in gpurun_out/r2m8_*.json; do echo $f; head -c 500 $f; echo; done; tail -5 gpurun_out/r2m8.err
