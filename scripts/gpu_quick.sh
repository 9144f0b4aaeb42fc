#!/bin/bash
# quick check on one GPU: a test selection + the default bench line
OUT=gpurun_out/${1:-quick}; mkdir -p $OUT
timeout 900 python -m pytest tests/test_gpu_level.py -k "valid_only or sedov_matches" -m gpu -x -q > $OUT/pytest.log 2>&1; tail -4 $OUT/pytest.log
timeout 300 python bench.py --steps 10 --warmup 3 > $OUT/bench.json 2>$OUT/bench.err; cat $OUT/bench.json | cut -c1-1600; tail -3 $OUT/bench.err
