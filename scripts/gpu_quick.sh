#!/bin/bash
# quick hydro check on one GPU: fused-path parity (exact + relaxed stage pairs, Sod/Sedov level tests) and both bench modes
OUT=gpurun_out/${1:-quick}; mkdir -p $OUT
timeout 900 python -m pytest tests/test_gpu_sweeps.py tests/test_gpu_relaxed.py tests/test_gpu_fullsize.py -m gpu -x -q > $OUT/pytest.log 2>&1; tail -4 $OUT/pytest.log
timeout 300 python bench.py --steps 10 --warmup 3 --no-extras > $OUT/bench_relaxed.json 2>$OUT/bench_relaxed.err; cat $OUT/bench_relaxed.json; tail -3 $OUT/bench_relaxed.err
