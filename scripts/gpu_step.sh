mkdir -p gpurun_out
timeout 300 ncu --set full --clock-control none --import-source on -k regex:conv_wgrad_tc_kernel -s 2 -c 1 -o gpurun_out/prof_wtc128 python scripts/profile_conv.py 128 128 32 4 wgrad 1 > gpurun_out/ncu_wtc128.log 2>&1
timeout 300 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err
wc -l gpurun_out/bench.json; cut -c1-300 gpurun_out/bench.json; tail -2 gpurun_out/bench.err
