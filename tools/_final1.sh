python -m pytest tests -m gpu -x -q > gpurun_out/gputests_r02.log 2>&1; tail -2 gpurun_out/gputests_r02.log
python bench.py > gpurun_out/bench_driver_r02.log 2>&1
python bench.py --steps 1000 --warmup 5 --e2e-steps 20 > gpurun_out/bench_plain_r02.log 2>&1
CRUX_PG_LIB=$PWD/variants/libpg_sync2.so python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/c3_sync2.log 2>&1
grep -o "\"ms_per_step\": [0-9.]*\|value_thermalised\": [0-9.]*\|\"e2e\": {\"value\": [0-9.]*" gpurun_out/bench_driver_r02.log gpurun_out/bench_plain_r02.log gpurun_out/c3_sync2.log
