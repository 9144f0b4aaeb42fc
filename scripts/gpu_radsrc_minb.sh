#!/bin/bash
# relaxed source-term solve: resident CTAs per SM the kernel is compiled for (register cap 64 / 80 / 96 / 128) inside config C4's coarse step
OUT=gpurun_out/${1:-r02_minb}; mkdir -p $OUT
for m in ${MINBS:-8 6 5 4}; do
QK_RADSRC_MINB=$m timeout 300 python bench.py --workload radhydro --arith relaxed --steps 2 --warmup 1 --no-extras --no-subrecords > $OUT/radhydro_minb$m.json 2> $OUT/radhydro_minb$m.err
python -c "
import json
d=json.loads(open('$OUT/radhydro_minb$m.json').read().strip().splitlines()[-1]); print('minb$m', d['value'], d['ms_per_step'], d['kernel_ms_per_step'].get('rad_source_terms'))" || tail -5 $OUT/radhydro_minb$m.err
done
