#!/bin/bash
# ncu --set full of one radiation substep of config C4 (relaxed arithmetic): k_rad_prim, k_rad_x, k_rad_m<Y>, k_rad_m<Z>, k_rad_source
# usage (under gpurun): bash scripts/gpu_ncu_rad.sh <tag>
TAG=${1:-r02_rad}
OUT=gpurun_out/$TAG
mkdir -p $OUT
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_rad_x|k_rad_m|k_rad_prim|k_rad_source|k_copy_comps' -s 110 -c 6 \
    -o $OUT/prof python bench.py --workload radhydro --arith relaxed --steps 1 --warmup 1 --no-extras > $OUT/ncu.log 2>&1
tail -3 $OUT/ncu.log
ncu -i $OUT/prof.ncu-rep --page raw --csv > $OUT/raw.csv 2>/dev/null
python scripts/ncu_summary.py $OUT/raw.csv > $OUT/summary.csv
ncu -i $OUT/prof.ncu-rep --page source --csv > $OUT/src.csv 2>/dev/null
for i in 0 1 2 3 4 5; do python scripts/ncu_opmix.py $OUT/src.csv 1 $i > $OUT/opmix_$i.txt 2>&1; done
rm -f $OUT/src.csv
ls -la $OUT
