mkdir -p gpurun_out/r01k
timeout 600 python -m pytest tests/test_gpu_radiation.py -m gpu -x -q > gpurun_out/r01k/pytest_rad.log 2>&1; tail -3 gpurun_out/r01k/pytest_rad.log
timeout 300 python bench.py --workload radiation --steps 10 --warmup 3 --no-extras > gpurun_out/r01k/bench_rad.json 2> gpurun_out/r01k/bench_rad.err; cat gpurun_out/r01k/bench_rad.json; tail -3 gpurun_out/r01k/bench_rad.err
timeout 300 python bench.py --steps 10 --warmup 3 --no-extras > gpurun_out/r01k/bench_relaxed.json 2>gpurun_out/r01k/bench_relaxed.err; cat gpurun_out/r01k/bench_relaxed.json
