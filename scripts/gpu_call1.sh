#!/bin/bash
# round-2 call 1: bisect the two masked failures, validate the opt-in kernels, default bench + layer trace
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/r2c1_smi.txt
timeout 400 python scripts/diag_toy.py > gpurun_out/diag_toy_default.txt 2>&1
for v in noside torchsgd nodirect mma; do
  timeout 120 python scripts/diag_toy.py --variant $v --no-oracle --checks 0,1,2,10,59 > gpurun_out/diag_toy_$v.txt 2>&1
done
timeout 200 python scripts/diag_inference.py > gpurun_out/diag_inference.txt 2>&1
: > gpurun_out/pytest_experimental.log
for t in test_strided_wgrad_on_tcgen05 test_transposed_conv_wgrad_on_tcgen05 test_strided_conv_block_forward_on_tcgen05 \
         test_upconv_input_gradient_on_tcgen05 test_item_order_repack_kernel test_bulk_weight_stream_in_the_tile_kernel \
         test_narrow_norm_backward_passes; do
  echo "== $t" >> gpurun_out/pytest_experimental.log
  NND_EXPERIMENTAL=1 timeout 200 python -m pytest tests/test_zz_experimental_gpu.py -q -x -k "$t" 2>&1 | tail -25 >> gpurun_out/pytest_experimental.log
done
timeout 300 python bench.py --no-cpu-baseline --trace-layers gpurun_out/layers.csv --steps 10 --warmup 3 > gpurun_out/r2c1_bench.json 2> gpurun_out/r2c1_bench.err
tail -3 gpurun_out/diag_toy_default.txt; grep -E '^==|passed|failed|error' gpurun_out/pytest_experimental.log; tail -5 gpurun_out/diag_inference.txt
