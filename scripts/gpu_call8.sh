#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu > gpurun_out/r2c8_pytest_gpu.log 2>&1; tail -6 gpurun_out/r2c8_pytest_gpu.log
timeout 300 python bench.py --no-cpu-baseline --steps 20 --warmup 5 --trace-layers gpurun_out/r2c8_layers.csv > gpurun_out/r2c8_bench.json 2> gpurun_out/r2c8_bench.err; head -c 300 gpurun_out/r2c8_bench.json; echo
timeout 300 python bench.py --no-cpu-baseline --steps 20 --warmup 5 > gpurun_out/r2c8_bench_b.json 2>> gpurun_out/r2c8_bench.err; head -c 300 gpurun_out/r2c8_bench_b.json; echo
timeout 300 python bench.py --no-cpu-baseline --steps 10 --warmup 3 --config adam > gpurun_out/r2c8_bench_adam.json 2>> gpurun_out/r2c8_bench.err; head -c 300 gpurun_out/r2c8_bench_adam.json; echo
timeout 300 python bench.py --no-cpu-baseline --steps 10 --warmup 3 --config lidc > gpurun_out/r2c8_bench_lidc.json 2>> gpurun_out/r2c8_bench.err; head -c 300 gpurun_out/r2c8_bench_lidc.json; echo
timeout 300 python scripts/bench_inference.py > gpurun_out/r2c8_infer1.json 2>> gpurun_out/r2c8_bench.err; cat gpurun_out/r2c8_infer1.json
tail -3 gpurun_out/r2c8_bench.err
