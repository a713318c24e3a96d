#!/bin/bash
# two-GPU diagnosis: why is the device-resident loop slower than the end-to-end loop at 2 GPUs (31.3 vs 21.8 ms)?
set -u
mkdir -p gpurun_out
run() { local tag=$1; shift
  timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $((29500 + RANDOM % 400)) \
     bench.py --gpus 2 --warmup 4 --no-cpu-baseline "$@" > gpurun_out/r2s2_2gpu_diag_$tag.json 2> gpurun_out/r2s2_2gpu_diag_$tag.err
  python -c "import json; d=json.load(open('gpurun_out/r2s2_2gpu_diag_$tag.json')); print('$tag: resident', round(d['ms_per_step'],2), 'e2e', round(d['e2e']['ms_per_step'],2))"; }
run default --steps 10
run no_buckets --steps 10 --experimental no_buckets
run old_kernels --steps 10 --experimental no_wgrad_tma,gather_tma=0
run no_wgrad_tma --steps 10 --experimental no_wgrad_tma
run gather0 --steps 10 --experimental gather_tma=0
run steps30 --steps 30
