#!/bin/bash
# N-GPU evidence run (usage under gpurun --gpus N: bash scripts/gpu_multi.sh <tag> N): the N-rank parity tests (hydro and config C4),
# the default bench line at N ranks (carries the `parity` sub-record: N ranks vs one rank, SHA-256), and config C4's coarse step at N ranks.
OUT=gpurun_out/${1:-r02_multi}; N=${2:-2}; mkdir -p $OUT
nvidia-smi --query-gpu=index,name --format=csv,noheader > $OUT/gpus.txt
timeout 1200 python -m pytest tests/test_gpu_multirank.py tests/test_zzzzgpu_multirank_shell.py -m gpu -q -rs > $OUT/pytest_multirank.log 2>&1; tail -12 $OUT/pytest_multirank.log
PORT=29541
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $PORT bench.py --gpus $N --steps 20 --warmup 5 > $OUT/bench_hydro_$N.json 2> $OUT/bench_hydro_$N.err
python - <<P
import json
try:
    d=json.loads(open('$OUT/bench_hydro_$N.json').read().strip().splitlines()[-1]); print('hydro N=$N', d['value'], d['ms_per_step'], d.get('parity'), d['e2e']['value'], d['kernel_ms_per_step'])
except Exception as e:
    print('hydro bench failed', e); print(open('$OUT/bench_hydro_$N.err').read()[-3000:])
P
for a in relaxed; do
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((PORT+1)) bench.py --workload radhydro --arith $a --gpus $N --steps 3 --warmup 1 --no-extras > $OUT/bench_radhydro_${a}_$N.json 2> $OUT/bench_radhydro_${a}_$N.err
python - <<P
import json
try:
    d=json.loads(open('$OUT/bench_radhydro_${a}_$N.json').read().strip().splitlines()[-1]); print('radhydro $a weak N=$N', d['value'], d['ms_per_step'], d['kernel_ms_per_step'])
except Exception as e:
    print('radhydro bench failed', e); print(open('$OUT/bench_radhydro_${a}_$N.err').read()[-3000:])
P
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((PORT+2)) bench.py --workload radhydro --arith $a --ncell 256 --gpus $N --steps 3 --warmup 1 --no-extras > $OUT/bench_radhydro_${a}_c4_$N.json 2> $OUT/bench_radhydro_${a}_c4_$N.err
python - <<P
import json
try:
    d=json.loads(open('$OUT/bench_radhydro_${a}_c4_$N.json').read().strip().splitlines()[-1]); print('radhydro $a 256^3 (configs[3]) N=$N', d['value'], d['ms_per_step'], d['kernel_ms_per_step'])
except Exception as e:
    print('radhydro c4 bench failed', e); print(open('$OUT/bench_radhydro_${a}_c4_$N.err').read()[-3000:])
P
done
