#!/bin/bash
# ncu --set full of the six sweep launches of one relaxed hydro step at 256^3 (usage under gpurun: bash scripts/gpu_ncu_hydro.sh <tag> [arith])
TAG=${1:-r02_hydro}; ARITH=${2:-relaxed}
OUT=gpurun_out/$TAG; mkdir -p $OUT
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_sweep_xc|k_sweep_xt|k_march_t' -s 36 -c 6 \
    -o $OUT/prof_$ARITH python bench.py --arith $ARITH --steps 2 --warmup 6 --no-extras > $OUT/ncu_$ARITH.log 2>&1
tail -3 $OUT/ncu_$ARITH.log
ncu -i $OUT/prof_$ARITH.ncu-rep --page raw --csv > $OUT/raw_$ARITH.csv 2>/dev/null
python scripts/ncu_summary.py $OUT/raw_$ARITH.csv > $OUT/summary_$ARITH.csv
ncu -i $OUT/prof_$ARITH.ncu-rep --page source --csv > $OUT/src.csv 2>/dev/null
for i in 0 1 2; do python scripts/ncu_opmix.py $OUT/src.csv 1 $i > $OUT/opmix_${ARITH}_$i.txt 2>&1; done
rm -f $OUT/src.csv
grep -E "Kernel Name|gpu__time_duration|dram__bytes|registers|warps_active|pipe_fp64|inst_issued|inst_executed.sum|stalled_(long|wait|math|short|not_sel|sleeping|mio|branch|no_inst|barrier|lg)" $OUT/summary_$ARITH.csv | cut -c1-150 | head -120
# the .ncu-rep (25-60 MB each) would push gpurun_out/ over the 64 MiB that travels back: keep the summaries only
rm -f $OUT/prof_$ARITH.ncu-rep $OUT/raw_$ARITH.csv
