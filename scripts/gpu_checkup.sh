#!/bin/bash
# One gpurun call that re-establishes the ground truth on a fresh B200 box (run from the repo root):
#   /usr/local/graft/bin/gpurun --timeout 2400 -- 'bash scripts/gpu_checkup.sh'      (typically ~15 min)
# Outputs land in gpurun_out/: pytest log (incl. XPASS/xfail of the tests written without a GPU), the default bench line (with the
# stock-PyTorch gpu_baseline and the CPU baseline), the per-launch convolution trace of one train step (layer geometry -> kernel -> ms),
# and the ncu launch list of the same step.  Every stage has its own timeout so a hang cannot take the box down with it.
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests -q -m gpu -rxX 2>&1 | tail -40 > gpurun_out/pytest_gpu.log
timeout 420 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err
timeout 240 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --trace-layers gpurun_out/layers.csv > gpurun_out/bench_trace.json 2>> gpurun_out/bench.err
# kernels written without a GPU (off by default): their tests in a process of their own, then an A/B bench line
# (one process per test function: an addressing bug in one kernel must not take the others' verdicts with it)
: > gpurun_out/pytest_experimental.log
for t in test_strided_wgrad_on_tcgen05 test_transposed_conv_wgrad_on_tcgen05 test_strided_conv_block_forward_on_tcgen05 \
         test_upconv_input_gradient_on_tcgen05 test_item_order_repack_kernel test_bulk_weight_stream_in_the_tile_kernel \
         test_narrow_norm_backward_passes; do
  echo "== $t" >> gpurun_out/pytest_experimental.log
  NND_EXPERIMENTAL=1 timeout 200 python -m pytest tests/test_zz_experimental_gpu.py -q -k "$t" 2>&1 | tail -6 >> gpurun_out/pytest_experimental.log
done
timeout 240 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --experimental wgrad_s2 > gpurun_out/bench_wgrad_s2.json 2>> gpurun_out/bench.err
timeout 240 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --experimental wgrad_s2,gather_s2 > gpurun_out/bench_s2_all.json 2>> gpurun_out/bench.err
timeout 240 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --experimental tc_bulk > gpurun_out/bench_tc_bulk.json 2>> gpurun_out/bench.err
timeout 240 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --experimental norm_narrow > gpurun_out/bench_norm_narrow.json 2>> gpurun_out/bench.err
timeout 420 ncu --metrics gpu__time_duration.sum --clock-control none -s 2150 -c 900 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 1 --warmup 3 --profile > gpurun_out/profile.log 2>&1
python - <<'PY'
import csv, collections
rows = list(csv.DictReader(open("gpurun_out/layers.csv")))
agg = collections.defaultdict(lambda: [0.0, 0.0, 0])
for r in rows:
    k = (r["kind"], r["kernel"], r["Cin"], r["Cout"], f'{r["Ld"]}x{r["Lh"]}x{r["Lw"]}', f's{r["sd"]}{r["sh"]}{r["sw"]}', r["T"])
    agg[k][0] += float(r["ms"]); agg[k][1] += float(r["gflop"]); agg[k][2] += 1
tot = sum(v[0] for v in agg.values())
with open("gpurun_out/layers_summary.txt", "w") as f:
    f.write(f"convolution-family launches of one train step: {len(rows)} launches, {tot:.2f} ms summed\n")
    f.write("ms      share%  n   TFLOP/s  kind         kernel         Cin->Cout  grid        stride taps\n")
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][0]):
        f.write(f"{v[0]:7.3f} {100 * v[0] / tot:6.2f} {v[2]:3d} {v[1] / max(v[0], 1e-9):8.1f}  {k[0]:12s} {k[1]:14s} {k[2]:>3s}->{k[3]:<4s} {k[4]:11s} {k[5]:6s} {k[6]}\n")
PY
tail -5 gpurun_out/pytest_gpu.log; grep -E '^==|passed|failed|error' gpurun_out/pytest_experimental.log; head -c 400 gpurun_out/bench_wgrad_s2.json; echo; head -c 400 gpurun_out/bench_s2_all.json; echo; head -c 400 gpurun_out/bench_tc_bulk.json; echo; head -c 400 gpurun_out/bench_norm_narrow.json; echo; cat gpurun_out/bench.json | head -c 1500; echo; head -25 gpurun_out/layers_summary.txt
# single-layer timings of the strided forms, default kernels vs the opt-in tcgen05 ones
for args in "32 64 128 4" "64 128 64 4" "128 256 32 4"; do
  for m in fprop wgrad; do
    NND_STRIDE=2 timeout 120 python scripts/profile_conv.py $args $m 2>&1 | tail -1
    NND_STRIDE=2 NND_S2=1 timeout 120 python scripts/profile_conv.py $args $m 2>&1 | tail -1
  done
done | tee gpurun_out/strided_layers.txt
# BASELINE config 4 through the predictor + ensembler (one 288^3 case, 160^3 patches, 8 mirror passes)
timeout 300 python scripts/bench_inference.py > gpurun_out/bench_inference.json 2>> gpurun_out/bench.err; cat gpurun_out/bench_inference.json
# CTA-pair probe (cta_group::2 semantics + issue rate) for the round-2 128-channel kernels; a hang only costs its timeout
if [ ! -x scripts/pair_probe ]; then nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I nndetection_b200/csrc -o scripts/pair_probe scripts/pair_probe.cu; fi
timeout 60 scripts/pair_probe 2>&1 | tee gpurun_out/pair_probe.txt
