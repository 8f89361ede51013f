python -m pytest tests/test_gpu_worlds.py tests/test_gpu_dropin.py tests/test_abi.py tests/test_debug_geometry.py -m gpu -x -q > gpurun_out/t_pre.log 2>&1; tail -2 gpurun_out/t_pre.log
python tests/random_campaign_mixed.py 150 9 > gpurun_out/campaign_pre.log 2>&1; tail -1 gpurun_out/campaign_pre.log
crux_b200/lib/crux_headless tests/golden/scenes/bouncehouse.json 3000 /tmp/o.bin --player | grep 'frame loop' > gpurun_out/c1_latency_pre.log
crux_b200/lib/crux_headless tests/golden/scenes/bouncehouse.json 6000 /tmp/o.bin --player | grep 'frame loop' >> gpurun_out/c1_latency_pre.log
PG_FRAME_TIMING=1 crux_b200/lib/crux_headless tests/golden/scenes/bouncehouse.json 3000 /tmp/o.bin --player 2>&1 | tail -4 >> gpurun_out/c1_latency_pre.log
cat gpurun_out/c1_latency_pre.log
