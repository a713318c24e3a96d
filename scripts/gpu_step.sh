mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_net_gpu.py -q -m gpu -k "all_taps or stacked_tap or single_layer or plain_conv or block_fwd" > gpurun_out/pytest_w.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_w.log
rm -f gpurun_out/conv_rates.txt
timeout 120 python scripts/profile_conv.py 128 128 32 4 wgrad 1 >> gpurun_out/conv_rates.txt 2>&1
timeout 120 python scripts/profile_conv.py 128 128 32 4 wgrad 3 >> gpurun_out/conv_rates.txt 2>&1
timeout 120 python scripts/profile_conv.py 128 128 16 4 wgrad 1 >> gpurun_out/conv_rates.txt 2>&1
timeout 120 python scripts/profile_conv.py 128 128 16 4 wgrad 3 >> gpurun_out/conv_rates.txt 2>&1
timeout 300 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err
tail -6 gpurun_out/pytest_w.log; cat gpurun_out/conv_rates.txt; cut -c1-330 gpurun_out/bench.json
