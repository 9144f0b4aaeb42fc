#!/bin/bash
# hydro sweeps after a staging change: bit-exact tests (fused vs faithful vs oracle, kept fluxes, level driver, relaxed drift) + the bench line
OUT=gpurun_out/${1:-tmap_check}; mkdir -p $OUT
timeout 600 python -m pytest tests/test_gpu_sweeps.py -m gpu -q -x > $OUT/pytest_sweeps.log 2>&1; tail -6 $OUT/pytest_sweeps.log
timeout 900 python -m pytest tests/test_gpu_keep_fluxes.py tests/test_gpu_relaxed.py tests/test_gpu_level.py -m gpu -q -s -x > $OUT/pytest.log 2>&1; tail -8 $OUT/pytest.log
for a in relaxed exact; do
timeout 300 python bench.py --arith $a --steps 20 --warmup 5 --no-extras > $OUT/bench_$a.json 2> $OUT/bench_$a.err
python -c "
import json
d=json.loads(open('$OUT/bench_$a.json').read().strip().splitlines()[-1]); print('$a', d['value'], d['ms_per_step'], d['kernel_ms_per_step'])" || tail -5 $OUT/bench_$a.err
done
