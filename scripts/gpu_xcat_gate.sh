#!/bin/bash
# the launch-size gate of the concatenated x sweep: its parity tests (forced on), the small AMR cases through the reference's driver (gate
# closed: per-row tiles) and the 256^3 bench line (gate open)
OUT=gpurun_out/${1:-r02_gate}; mkdir -p $OUT
timeout 300 python -m pytest tests/test_gpu_sweeps.py -m gpu -q -x > $OUT/pytest_sweeps.log 2>&1; tail -3 $OUT/pytest_sweeps.log
timeout 300 python scripts/gpu_refcuda.py --only sedov_amr64_maxlev2,sod_c1_amr,sedov_amr256_c5 --out $OUT/ref_cuda_amr.json > $OUT/ref_cuda_amr.log 2>&1
grep -E "^(sedov|sod).* (stock|exact|relaxed) [0-9]|CPU reference bits" $OUT/ref_cuda_amr.log | cut -c1-220
timeout 200 python bench.py --steps 20 --warmup 5 --no-extras --no-subrecords > $OUT/bench_relaxed.json 2> $OUT/bench_relaxed.err
python -c "
import json
d=json.loads(open('$OUT/bench_relaxed.json').read().strip().splitlines()[-1]); print('relaxed', d['value'], d['ms_per_step'], d['kernel_ms_per_step'], d['roofline']['traffic'])"
