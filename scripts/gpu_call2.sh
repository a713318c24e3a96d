#!/bin/bash
set -u
mkdir -p gpurun_out
for i in 1 2 3; do timeout 300 python -m pytest tests/test_zz_inference_gpu.py -q 2>&1 | tail -40 > gpurun_out/r2c2_inference_$i.log; done
timeout 900 python -m pytest tests/test_zz_toy_training_gpu.py tests/test_zz_fullsize_parity_gpu.py tests/test_nms_gpu.py tests/test_zz_trace_gpu.py -q 2>&1 | tail -80 > gpurun_out/r2c2_newtests.log
for e in "" wgrad_s2 wgrad_s2,gather_s2 norm_narrow wgrad_s2,gather_s2,norm_narrow; do
  n=$(echo "$e" | tr ',' '_'); [ -z "$n" ] && n=default
  timeout 240 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --experimental "$e" > gpurun_out/r2c2_bench_$n.json 2>> gpurun_out/r2c2_bench.err
done
timeout 120 python - > gpurun_out/r2c2_refnms.json 2>> gpurun_out/r2c2_bench.err <<'PY'
import json, sys
sys.path.insert(0, ".")
import torch, bench
print(json.dumps(bench.ref_nms_rates("cuda:0"), indent=1))
PY
tail -3 gpurun_out/r2c2_inference_*.log; tail -15 gpurun_out/r2c2_newtests.log; for f in gpurun_out/r2c2_bench_*.json; do echo $f; head -c 300 $f; echo; done; cat gpurun_out/r2c2_refnms.json
