#!/bin/bash
# round 2, session 2, call 8: first device run of the strided TMA wgrad (tests under a timeout), toy test with the new gate, step + trace
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_wgrad_tma_gpu.py tests/test_strided_tcgen05_gpu.py tests/test_zz_toy_training_gpu.py -q -m gpu --timeout 120 2>&1 | tail -40 | tee gpurun_out/r2s2c8_tests.txt
timeout 300 python bench.py --no-cpu-baseline --steps 20 --warmup 5 --trace-layers gpurun_out/r2s2c8_layers.csv > gpurun_out/r2s2c8_bench.json 2> gpurun_out/r2s2c8_bench.err; head -c 300 gpurun_out/r2s2c8_bench.json; echo
python scripts/timeline.py gpurun_out/r2s2c8_layers.csv | grep "wgrad_tma_s2" | head -20
