set -x
CMD="python bench.py --steps 20 --warmup 5 --no-cpu-baseline --thermalise-steps 0 --e2e-steps 3"
python bench.py --steps 20 --warmup 5 > gpurun_out/bench_driver_r02.log 2>&1
python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/bench_ref_r02.log 2>&1
$CMD > gpurun_out/bench_short_r02.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r02.csv $CMD > gpurun_out/ncu_launches.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_step_spheres -s 2 -c 2 -o gpurun_out/k_step_spheres_r02 -f $CMD > gpurun_out/ncu_full.log 2>&1
tail -c 400 gpurun_out/bench_driver_r02.log; tail -c 300 gpurun_out/bench_ref_r02.log
