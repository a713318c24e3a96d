#!/bin/bash
# two-GPU check of the session-2 build: default step (gradient buckets during backward) + the bucket-vs-flat gradient comparison
set -u
mkdir -p gpurun_out
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $((29500 + RANDOM % 400)) \
   bench.py --gpus 2 --steps 10 --warmup 4 --no-cpu-baseline > gpurun_out/r2s2_2gpu_bench.json 2> gpurun_out/r2s2_2gpu_bench.err
python -c "import json; d=json.load(open('gpurun_out/r2s2_2gpu_bench.json')); print('2 GPUs:', d['ms_per_step'], d['value'], d['e2e']['ms_per_step'], d['n_gpus'])"
tail -3 gpurun_out/r2s2_2gpu_bench.err
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $((29500 + RANDOM % 400)) \
   scripts/ddp_bucket_check.py 2>&1 | tail -6 | tee gpurun_out/r2s2_2gpu_bucket_check.txt
