#!/bin/bash
# round 2, session 2, last call: the TMA-fed 32-channel stacked-tap wgrad -- its tests, the full-size train step, the learning check, the step time
set -u
mkdir -p gpurun_out
timeout 150 python -m pytest tests/test_wgrad_tma_gpu.py "tests/test_net_gpu.py::test_full_size_configs_properties" tests/test_zz_toy_training_gpu.py -q -m gpu --timeout 90 2>&1 | tail -5 | tee gpurun_out/r2s2l_tests.txt
timeout 120 python bench.py --no-cpu-baseline --steps 20 --warmup 5 > gpurun_out/r2s2l_bench.json 2> gpurun_out/r2s2l_bench.err
python -c "
import json; d=json.load(open('gpurun_out/r2s2l_bench.json')); print('bench:', round(d['ms_per_step'],2), 'ms', round(d['value'],1), 'patches/s; e2e', round(d['e2e']['ms_per_step'],2))
for r in d['roofline_kernels']:
    if '32->32' in r.get('layer',''): print(r)"
