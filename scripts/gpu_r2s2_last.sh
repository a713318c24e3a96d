#!/bin/bash
# round 2, session 2, last call: the TMA-fed 32-channel stacked-tap wgrad -- its tests, the whole suite, the step
set -u
mkdir -p gpurun_out
timeout 120 python -m pytest tests/test_wgrad_tma_gpu.py -q -m gpu --timeout 60 -k "32_channel" 2>&1 | tail -4 | tee gpurun_out/r2s2l_tests32.txt
timeout 400 python -m pytest tests -q -m gpu --timeout 300 -x > gpurun_out/r2s2l_pytest_gpu.log 2>&1; tail -3 gpurun_out/r2s2l_pytest_gpu.log
timeout 200 python bench.py --no-cpu-baseline --steps 20 --warmup 5 > gpurun_out/r2s2l_bench.json 2> gpurun_out/r2s2l_bench.err
python -c "
import json; d=json.load(open('gpurun_out/r2s2l_bench.json')); print('bench:', round(d['ms_per_step'],2), 'ms', round(d['value'],1), 'patches/s; e2e', round(d['e2e']['ms_per_step'],2))
for r in d['roofline_kernels']:
    if '32->32' in r.get('layer',''): print(r)"
