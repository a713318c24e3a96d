#!/bin/bash
# round 2, session 2, call 5: the failing full-size test in full, the learning check, per-stream timeline of one step
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest "tests/test_net_gpu.py::test_full_size_configs_properties" -x -q -m gpu 2>&1 | tail -60 > gpurun_out/r2s2c5_fullsize.txt; tail -30 gpurun_out/r2s2c5_fullsize.txt
timeout 600 python -m pytest tests/test_zz_toy_training_gpu.py tests/test_zz_fullsize_parity_gpu.py -x -q -m gpu 2>&1 | tail -15 | tee gpurun_out/r2s2c5_toy.txt
timeout 300 python bench.py --no-cpu-baseline --steps 10 --warmup 5 --trace-layers gpurun_out/r2s2c5_layers.csv > gpurun_out/r2s2c5_bench.json 2> gpurun_out/r2s2c5_bench.err; head -c 300 gpurun_out/r2s2c5_bench.json; echo
python scripts/timeline.py gpurun_out/r2s2c5_layers.csv > gpurun_out/r2s2c5_timeline.txt; head -5 gpurun_out/r2s2c5_timeline.txt
