mkdir -p gpurun_out/r01i
timeout 600 python bench.py --workload radiation --steps 10 --warmup 3 > gpurun_out/r01i/bench_rad.json 2> gpurun_out/r01i/bench_rad.err; cat gpurun_out/r01i/bench_rad.json; tail -5 gpurun_out/r01i/bench_rad.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'k_rad_stage|k_rad_prim' -s 8 -c 4 -o gpurun_out/r01i/prof_rad python bench.py --workload radiation --steps 2 --warmup 2 --no-extras > gpurun_out/r01i/ncu_rad.log 2>&1; tail -3 gpurun_out/r01i/ncu_rad.log
timeout 300 python bench.py --steps 10 --warmup 3 --no-extras > gpurun_out/r01i/bench_relaxed.json 2>gpurun_out/r01i/bench_relaxed.err; cat gpurun_out/r01i/bench_relaxed.json
