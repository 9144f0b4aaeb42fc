#!/bin/bash
# ncu launch list (gpu__time_duration.sum per launch) of the default bench command, both arithmetic modes; aggregated per kernel
OUT=gpurun_out/${1:-r02_launches}; mkdir -p $OUT
for a in relaxed exact; do
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 30 -c 400 --csv --log-file $OUT/launches_$a.csv python bench.py --arith $a --steps 4 --warmup 2 --no-extras --no-subrecords > $OUT/bench_under_ncu_$a.log 2>&1
( echo "# ncu launch list of bench.py --arith $a --steps 4 --warmup 2 --no-extras --no-subrecords (final kernels of round 2), Sedov 256^3, B200, unit ns (cold-cache, serialised: compare SHARES with bench.py kernel_ms_per_step)"
  echo "# command: ncu --metrics gpu__time_duration.sum --clock-control none -s 30 -c 400 --csv --log-file launches.csv python bench.py ..."
  python scripts/launch_summary.py $OUT/launches_$a.csv ) > $OUT/summary_$a.csv
head -14 $OUT/summary_$a.csv
rm -f $OUT/launches_$a.csv
done
