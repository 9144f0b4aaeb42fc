#!/bin/bash
# the concatenated x sweep (k_sweep_xc) against the per-row tiles (QK_XCAT=0): bit-exact tests + bench lines in both modes
OUT=gpurun_out/${1:-xcat}; mkdir -p $OUT
timeout 900 python -m pytest tests/test_gpu_sweeps.py -m gpu -q -x > $OUT/pytest_sweeps.log 2>&1; tail -6 $OUT/pytest_sweeps.log
timeout 900 python -m pytest tests/test_gpu_keep_fluxes.py tests/test_gpu_relaxed.py tests/test_gpu_level.py -m gpu -q -x > $OUT/pytest.log 2>&1; tail -8 $OUT/pytest.log
for x in 1 0; do
for a in relaxed exact; do
QK_XCAT=$x timeout 300 python bench.py --arith $a --steps 20 --warmup 5 --no-extras --no-subrecords > $OUT/bench_${a}_xcat$x.json 2> $OUT/bench_${a}_xcat$x.err
python -c "
import json
d=json.loads(open('$OUT/bench_${a}_xcat$x.json').read().strip().splitlines()[-1]); print('xcat$x $a', d['value'], d['ms_per_step'], d['kernel_ms_per_step'])" || tail -5 $OUT/bench_${a}_xcat$x.err
done
done
