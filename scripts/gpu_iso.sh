#!/bin/bash
# isothermal-EOS parity tests + the hydro suites that the chi early-out touches + bench lines in both modes
OUT=gpurun_out/${1:-r02_iso}; mkdir -p $OUT
timeout 900 python -m pytest tests/test_gpu_operators.py tests/test_gpu_level.py -m gpu -q -x > $OUT/pytest_iso.log 2>&1; tail -6 $OUT/pytest_iso.log
timeout 900 python -m pytest tests/test_gpu_sweeps.py tests/test_gpu_relaxed.py tests/test_gpu_keep_fluxes.py -m gpu -q -x > $OUT/pytest_hydro.log 2>&1; tail -6 $OUT/pytest_hydro.log
for a in relaxed exact; do
timeout 300 python bench.py --arith $a --steps 20 --warmup 5 --no-extras --no-subrecords > $OUT/bench_$a.json 2> $OUT/bench_$a.err
python -c "
import json
d=json.loads(open('$OUT/bench_$a.json').read().strip().splitlines()[-1]); print('$a', d['value'], d['ms_per_step'], d['kernel_ms_per_step'])" || tail -5 $OUT/bench_$a.err
done
