#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_pointwise_tma_gpu.py -q -x > gpurun_out/r2c4_pw_tests.log 2>&1
tail -25 gpurun_out/r2c4_pw_tests.log
timeout 300 python -m pytest tests/test_zz_inference_gpu.py tests/test_net_gpu.py -q -k "inference or plain_conv or transposed" > gpurun_out/r2c4_other_tests.log 2>&1
tail -5 gpurun_out/r2c4_other_tests.log
timeout 300 python bench.py --no-cpu-baseline --steps 10 --warmup 3 --trace-layers gpurun_out/r2c4_layers.csv > gpurun_out/r2c4_bench.json 2> gpurun_out/r2c4_bench.err
timeout 300 python bench.py --no-cpu-baseline --steps 10 --warmup 3 --experimental pw_tma --trace-layers gpurun_out/r2c4_layers_pw.csv > gpurun_out/r2c4_bench_pw.json 2>> gpurun_out/r2c4_bench.err
head -c 400 gpurun_out/r2c4_bench.json; echo; head -c 400 gpurun_out/r2c4_bench_pw.json; echo; tail -3 gpurun_out/r2c4_bench.err
