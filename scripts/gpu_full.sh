#!/bin/bash
# what the driver runs at round end, on one GPU: the whole -m gpu suite, smoke(), the default bench line and the reference arm
OUT=gpurun_out/${1:-full}; mkdir -p $OUT
( time timeout 2400 python -m pytest tests/ -x -q -m gpu > $OUT/pytest_gpu.log 2>&1 ) 2> $OUT/pytest_time.log; tail -6 $OUT/pytest_gpu.log; tail -3 $OUT/pytest_time.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/smoke.log 2>&1; tail -2 $OUT/smoke.log
( time timeout 1200 python bench.py > $OUT/bench_default.json 2> $OUT/bench_default.err ) 2> $OUT/bench_time.log; tail -3 $OUT/bench_time.log
python - <<P
import json
try:
    d=json.loads(open('$OUT/bench_default.json').read().strip().splitlines()[-1])
    print('value', d['value'], 'ms', d['ms_per_step'], 'e2e', d['e2e']['value'], 'roof', d['roofline'].get('frac'), d['roofline'].get('kernel'), 'traffic', d['roofline'].get('traffic'))
    print('clocks', d['clocks'])
    print('exact', d.get('exact_arith',{}).get('value'))
    print('north', {k:d.get('north_star_512',{}).get(k) for k in ('value','ms_per_step','sweeps','error','skipped')})
    print('radhydro', {k:d.get('radhydro',{}).get(k) for k in ('value','ms_per_step','error')})
    print('gpu_ref', json.dumps(d.get('gpu_reference'))[:600])
    print('cpu', d.get('cpu_baseline'))
except Exception as e:
    print('bench failed', e); print(open('$OUT/bench_default.err').read()[-3000:])
P
