#!/bin/bash
# relaxed hydro kernels after a change: drift tests + the bench line without extras (usage under gpurun: bash scripts/gpu_relaxed_check.sh <tag>)
OUT=gpurun_out/${1:-relaxed_check}; mkdir -p $OUT
timeout 900 python -m pytest tests/test_gpu_relaxed.py tests/test_gpu_sweeps.py tests/test_gpu_keep_fluxes.py -m gpu -q -s -x > $OUT/pytest.log 2>&1; tail -12 $OUT/pytest.log
for a in relaxed exact; do
timeout 300 python bench.py --arith $a --steps 20 --warmup 5 --no-extras > $OUT/bench_$a.json 2> $OUT/bench_$a.err
python -c "
import json
d=json.loads(open('$OUT/bench_$a.json').read().strip().splitlines()[-1]); print('$a', d['value'], d['ms_per_step'], d['kernel_ms_per_step'])" || tail -5 $OUT/bench_$a.err
done
