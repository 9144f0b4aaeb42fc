#!/bin/bash
# quick check on N GPUs: a test selection + the default bench line (usage under gpurun [--gpus N]: bash scripts/gpu_quick.sh <tag> [N] [pytest -k expr])
OUT=gpurun_out/${1:-quick}; N=${2:-1}; K=${3:-"lower_order or fused_equals or multirank"}; mkdir -p $OUT
timeout 900 python -m pytest tests -k "$K" -m gpu -x -q > $OUT/pytest.log 2>&1; tail -4 $OUT/pytest.log
if [ "$N" -gt 1 ]; then
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps 10 --warmup 3 --no-extras > $OUT/bench$N.json 2>$OUT/bench$N.err
else
  timeout 300 python bench.py --steps 10 --warmup 3 --no-extras > $OUT/bench$N.json 2>$OUT/bench$N.err
fi
python -c "
import json,sys
d=json.loads(open('$OUT/bench$N.json').read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], d['kernel_ms_per_step'])"; tail -2 $OUT/bench$N.err
