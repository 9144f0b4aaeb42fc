#!/bin/bash
# radiation path on one GPU: parity tests + bench (usage under gpurun: bash scripts/gpu_rad.sh <tag>)
OUT=gpurun_out/${1:-rad}; mkdir -p $OUT
timeout 600 python -m pytest tests/test_gpu_radiation.py -m gpu -x -q > $OUT/pytest_rad.log 2>&1; tail -5 $OUT/pytest_rad.log
timeout 300 python bench.py --workload radiation --steps 10 --warmup 3 --no-extras > $OUT/bench_rad.json 2> $OUT/bench_rad.err; cat $OUT/bench_rad.json; tail -3 $OUT/bench_rad.err
QK_RAD_TILE=1 timeout 300 python bench.py --workload radiation --steps 5 --warmup 2 --no-extras > $OUT/bench_rad_tile.json 2> $OUT/bench_rad_tile.err; cat $OUT/bench_rad_tile.json
