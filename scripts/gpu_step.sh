mkdir -p gpurun_out
timeout 600 python -m pytest tests -q -m gpu > gpurun_out/pytest_gpu.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_gpu.log
rm -f gpurun_out/conv_rates.txt
timeout 120 python scripts/profile_conv.py 32 32 128 4 wgrad 1 >> gpurun_out/conv_rates.txt 2>&1
timeout 300 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err
tail -5 gpurun_out/pytest_gpu.log; cat gpurun_out/conv_rates.txt; cat gpurun_out/bench.json
