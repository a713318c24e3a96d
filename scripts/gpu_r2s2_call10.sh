#!/bin/bash
# round 2, session 2, call 10: whole GPU suite + smoke + the other BASELINE configs with the TMA-fed kernels default
set -u
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -q -m gpu --timeout 600 > gpurun_out/r2s2c10_pytest_gpu.log 2>&1; tail -6 gpurun_out/r2s2c10_pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
for c in luna lidc adam; do
  timeout 400 python bench.py --no-cpu-baseline --config $c --steps 10 --warmup 4 > gpurun_out/r2s2c10_bench_$c.json 2> gpurun_out/r2s2c10_bench_$c.err
  echo "$c: $(head -c 400 gpurun_out/r2s2c10_bench_$c.json | grep -o '"value": [0-9.]*, "unit": "patches/s", "n_gpus": 1, "steps": 10, "warmup": 4, "ms_per_step": [0-9.]*')"
done
