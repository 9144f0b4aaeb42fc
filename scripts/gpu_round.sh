#!/bin/bash
# one GPU visit: parity tests, bench line, ncu launch list + full capture of the sweep kernels
# usage (under gpurun): bash scripts/gpu_round.sh <tag> [skip_tests]
TAG=${1:-run}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $OUT/smi.txt 2>&1
nproc > $OUT/nproc.txt
if [ -z "$2" ]; then
  timeout 1500 python -m pytest tests -m gpu -x -q > $OUT/pytest.log 2>&1; echo "pytest rc=$?" >> $OUT/pytest.log
  tail -5 $OUT/pytest.log
fi
timeout 900 python bench.py --steps 10 --warmup 3 > $OUT/bench.json 2> $OUT/bench.err; echo "bench rc=$?"
cat $OUT/bench.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 20 -c 400 --csv --log-file $OUT/launches.csv python bench.py --steps 2 --warmup 1 --no-extras > $OUT/ncu_launch.log 2>&1
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:'k_march_t|k_sweep_xt' -s 6 -c 6 -o $OUT/prof_sweeps python bench.py --steps 2 --warmup 1 --no-extras > $OUT/ncu_full.log 2>&1
ls -la $OUT
timeout 300 python __graft_entry__.py smoke > $OUT/smoke.log 2>&1; tail -2 $OUT/smoke.log
