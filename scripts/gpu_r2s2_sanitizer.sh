#!/bin/bash
# compute-sanitizer memcheck over smoke() with the TMA-fed kernels default (tiny train step: conv_tct / conv_tct S2 / wgrad_tma (paired rows) / wgrad_tma S2)
set -u
mkdir -p gpurun_out
timeout 170 compute-sanitizer --tool memcheck --print-limit 20 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2s2_sanitizer.log 2>&1; tail -6 gpurun_out/r2s2_sanitizer.log
