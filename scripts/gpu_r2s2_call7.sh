#!/bin/bash
# round 2, session 2, call 7: whole GPU suite with the TMA kernels on by default (toy learning gate aside), toy spread details, bench + trace
set -u
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu --timeout 600 --deselect tests/test_zz_toy_training_gpu.py 2>&1 | tail -15 | tee gpurun_out/r2s2c7_tests.txt
timeout 300 python scripts/calib_toy.py 2>&1 | grep "tma=" | tee gpurun_out/r2s2c7_calib_toy.txt
timeout 300 python bench.py --no-cpu-baseline --steps 20 --warmup 5 --trace-layers gpurun_out/r2s2c7_layers.csv > gpurun_out/r2s2c7_bench.json 2> gpurun_out/r2s2c7_bench.err; head -c 300 gpurun_out/r2s2c7_bench.json; echo
