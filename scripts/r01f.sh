mkdir -p gpurun_out/r01f
timeout 1200 python -m pytest tests/test_gpu_relaxed.py tests/test_gpu_level.py tests/test_gpu_sweeps.py -m gpu -x -q -s > gpurun_out/r01f/pytest.log 2>&1; tail -12 gpurun_out/r01f/pytest.log
timeout 300 python bench.py --steps 10 --warmup 3 --no-extras > gpurun_out/r01f/bench_exact.json 2>gpurun_out/r01f/bench_exact.err; cat gpurun_out/r01f/bench_exact.json
timeout 300 python bench.py --steps 10 --warmup 3 --no-extras --arith relaxed > gpurun_out/r01f/bench_relaxed.json 2>gpurun_out/r01f/bench_relaxed.err; cat gpurun_out/r01f/bench_relaxed.json; tail -3 gpurun_out/r01f/bench_relaxed.err
