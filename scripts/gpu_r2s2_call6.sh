#!/bin/bash
# round 2, session 2, call 6: toy-task spread (gate calibration), first device run of the TMA tile kernel (tests under a timeout), step A/B
set -u
mkdir -p gpurun_out
timeout 300 python scripts/calib_toy.py 2>&1 | grep "tma=" | tee gpurun_out/r2s2c6_calib_toy.txt
timeout 900 python -m pytest tests/test_gather_tma_gpu.py -q -m gpu --timeout 90 2>&1 | tail -40 | tee gpurun_out/r2s2c6_tests.txt
for m in 0 1 3; do
  timeout 300 python bench.py --no-cpu-baseline --steps 10 --warmup 5 --experimental gather_tma=$m --trace-layers gpurun_out/r2s2c6_layers_m$m.csv > gpurun_out/r2s2c6_bench_m$m.json 2> gpurun_out/r2s2c6_bench_m$m.err
  echo "gather_tma=$m: $(head -c 260 gpurun_out/r2s2c6_bench_m$m.json | grep -o '"ms_per_step": [0-9.]*')"
done
