#!/bin/bash
# round 2, session 2: final one-GPU check (whole GPU suite, smoke(), default bench line without the CPU legs)
set -u
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -q -m gpu --timeout 600 > gpurun_out/r2s2f_pytest_gpu.log 2>&1; tail -4 gpurun_out/r2s2f_pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 400 python bench.py --no-cpu-baseline --steps 20 --warmup 5 > gpurun_out/r2s2f_bench.json 2> gpurun_out/r2s2f_bench.err
python -c "import json; d=json.load(open('gpurun_out/r2s2f_bench.json')); print('bench:', round(d['ms_per_step'],2), 'ms', round(d['value'],1), 'patches/s; e2e', round(d['e2e']['ms_per_step'],2), 'ms; launches', d['gpu_launches'])"
