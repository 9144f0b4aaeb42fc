mkdir -p gpurun_out/r01e
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_march_t|k_sweep_xt' -s 6 -c 6 -o gpurun_out/r01e/prof_relaxed python bench.py --steps 2 --warmup 1 --no-extras --arith relaxed > gpurun_out/r01e/ncu_full.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 30 -c 200 --csv --log-file gpurun_out/r01e/launches.csv python bench.py --steps 2 --warmup 1 --no-extras --arith relaxed > gpurun_out/r01e/ncu_launch.log 2>&1
ls -la gpurun_out/r01e
