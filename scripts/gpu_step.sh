mkdir -p gpurun_out
./scripts/mma_rate > gpurun_out/mma_rate2.txt 2>&1
timeout 300 python -m pytest tests/test_net_gpu.py -q -k "streaming or single_layer" > gpurun_out/pytest_tcs.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_tcs.log
for cfg in "32 32 128 4" "64 64 64 4"; do
  for st in "0,2" "1,1" "1,2"; do
    NND_STREAM=$st timeout 120 python scripts/profile_conv.py $cfg fprop >> gpurun_out/conv_rates.txt 2>&1
  done
done
tail -5 gpurun_out/pytest_tcs.log; cat gpurun_out/conv_rates.txt; head -30 gpurun_out/mma_rate2.txt
