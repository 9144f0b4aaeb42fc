#!/bin/bash
# short N-GPU check of the final kernels (usage under gpurun --gpus N: bash scripts/gpu_multi_quick.sh <tag> N): the N-rank parity tests
# (hydro and config C4) and the default bench line at N ranks exactly as the driver launches it (carries the `parity` sub-record)
OUT=gpurun_out/${1:-r02_multiq}; N=${2:-2}; mkdir -p $OUT
timeout 600 python -m pytest tests/test_gpu_multirank.py tests/test_zzzzgpu_multirank_shell.py -m gpu -q -rs > $OUT/pytest_multirank.log 2>&1; tail -8 $OUT/pytest_multirank.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus $N --steps 20 --warmup 5 > $OUT/bench_hydro_$N.json 2> $OUT/bench_hydro_$N.err
python - <<P
import json
try:
    d=json.loads(open('$OUT/bench_hydro_$N.json').read().strip().splitlines()[-1]); print('hydro N=$N', d['value'], d['ms_per_step'], d.get('parity'), d['e2e']['value'], d['kernel_ms_per_step'])
except Exception as e:
    print('hydro bench failed', e); print(open('$OUT/bench_hydro_$N.err').read()[-3000:])
P
