#!/bin/bash
# round 2, session 2, call 3: TMA wgrad after the incremental-cursor fix: tests, component timing, step A/B
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_wgrad_tma_gpu.py tests/test_net_gpu.py -x -q -m gpu 2>&1 | tail -5 | tee gpurun_out/r2s2c3_tests.txt
{
timeout 120 python scripts/time_wgrad_tma.py 128 128 32 4 0 1 65 9 25 41 2>&1 | grep wgrad
timeout 120 python scripts/time_wgrad_tma.py 64 64 64 4 0 1 65 9 25 41 2>&1 | grep wgrad
timeout 120 python scripts/time_wgrad_tma.py 256 256 16 4 0 1 65 9 2>&1 | grep wgrad
timeout 120 python scripts/time_wgrad_tma.py 128 128 16 4 0 1 65 9 2>&1 | grep wgrad
timeout 120 python scripts/time_wgrad_tma.py 128 128 8 4 0 1 65 9 2>&1 | grep wgrad
} | tee gpurun_out/r2s2c3_components.txt
timeout 300 python bench.py --no-cpu-baseline --steps 20 --warmup 5 > gpurun_out/r2s2c3_bench_tma.json 2> gpurun_out/r2s2c3_bench_tma.err; head -c 400 gpurun_out/r2s2c3_bench_tma.json; echo
timeout 300 python bench.py --no-cpu-baseline --steps 20 --warmup 5 --experimental no_wgrad_tma > gpurun_out/r2s2c3_bench_notma.json 2> gpurun_out/r2s2c3_bench_notma.err; head -c 400 gpurun_out/r2s2c3_bench_notma.json; echo
