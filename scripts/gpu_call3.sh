#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu -x --durations=15 > gpurun_out/r2c3_pytest_gpu.log 2>&1
tail -30 gpurun_out/r2c3_pytest_gpu.log
timeout 400 python bench.py --no-cpu-baseline --steps 20 --warmup 5 > gpurun_out/r2c3_bench.json 2> gpurun_out/r2c3_bench.err
timeout 240 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --trace-layers gpurun_out/r2c3_layers.csv > /dev/null 2>> gpurun_out/r2c3_bench.err
timeout 420 ncu --metrics gpu__time_duration.sum --clock-control none -s 2150 -c 900 --csv --log-file gpurun_out/r2c3_launches.csv \
    python bench.py --steps 1 --warmup 3 --profile > gpurun_out/r2c3_profile.log 2>&1
head -c 600 gpurun_out/r2c3_bench.json; echo
