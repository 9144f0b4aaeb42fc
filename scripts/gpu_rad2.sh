#!/bin/bash
# round-2 radiation check on one GPU: parity tests of the transport sweeps (TMA-staged exact, first-generation, relaxed), C4 parity tests,
# transport bench in both arithmetic modes and both reconstruction orders, C4 bench in both modes.  usage under gpurun: bash scripts/gpu_rad2.sh <tag>
OUT=gpurun_out/${1:-r02_rad2}; mkdir -p $OUT
timeout 900 python -m pytest tests/test_gpu_radiation.py tests/test_zgpu_shell.py tests/test_zzgpu_relaxed_radhydro.py -m gpu -q -s > $OUT/pytest_rad.log 2>&1; tail -15 $OUT/pytest_rad.log
for a in exact relaxed; do for o in 3 2; do
  QK_BENCH_RAD_ORDER=$o timeout 300 python bench.py --workload radiation --arith $a --steps 10 --warmup 3 --no-extras > $OUT/bench_rad_${a}_o$o.json 2> $OUT/bench_rad_${a}_o$o.err
  python - <<P
import json
d=json.loads(open('$OUT/bench_rad_${a}_o$o.json').read().strip().splitlines()[-1]); print('radiation $a order $o', d['value'], d['ms_per_step'], d['kernel_ms_per_step'])
P
done; done
QK_RAD_V1=1 timeout 300 python bench.py --workload radiation --arith exact --steps 5 --warmup 2 --no-extras > $OUT/bench_rad_v1.json 2> $OUT/bench_rad_v1.err
python -c "
import json
d=json.loads(open('$OUT/bench_rad_v1.json').read().strip().splitlines()[-1]); print('radiation V1', d['value'], d['ms_per_step'], d['kernel_ms_per_step'])"
for a in exact relaxed; do
  timeout 300 python bench.py --workload radhydro --arith $a --steps 3 --warmup 1 --no-extras > $OUT/bench_radhydro_$a.json 2> $OUT/bench_radhydro_$a.err
  python -c "
import json
d=json.loads(open('$OUT/bench_radhydro_$a.json').read().strip().splitlines()[-1]); print('radhydro $a', d['value'], d['ms_per_step'], d['kernel_ms_per_step'])"
done
