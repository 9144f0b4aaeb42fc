mkdir -p gpurun_out/r01h
timeout 600 python -m pytest tests/test_gpu_level.py tests/test_gpu_sweeps.py -m gpu -x -q > gpurun_out/r01h/pytest.log 2>&1; tail -4 gpurun_out/r01h/pytest.log
timeout 300 python bench.py --steps 10 --warmup 3 --no-extras > gpurun_out/r01h/bench_relaxed.json 2>gpurun_out/r01h/bench_relaxed.err; cat gpurun_out/r01h/bench_relaxed.json; tail -3 gpurun_out/r01h/bench_relaxed.err
timeout 300 python bench.py --steps 10 --warmup 3 --no-extras --arith exact > gpurun_out/r01h/bench_exact.json 2>gpurun_out/r01h/bench_exact.err; cat gpurun_out/r01h/bench_exact.json
