#!/bin/bash
# ncu evidence for round 1 (run under gpurun, 1 GPU).  Each ncu pass is preceded by the identical
# plain command, which must exit 0 first.
set -u
OUT=gpurun_out
mkdir -p $OUT
B32="python bench.py --steps 2 --warmup 1 --batch 32 --sweep= --no-cpu-baseline"
B4K="python bench.py --steps 2 --warmup 1 --batch 4096 --sweep= --no-cpu-baseline"
# launch list (every launch with its device time)
$B32 > $OUT/plain_b32.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -s 150 -c 400 --csv --log-file $OUT/launches_b32.csv $B32 > $OUT/ncu_launches_b32.log 2>&1
echo "launch list b32 rc=$?"
$B4K > $OUT/plain_b4096.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -s 150 -c 400 --csv --log-file $OUT/launches_b4096.csv $B4K > $OUT/ncu_launches_b4096.log 2>&1
echo "launch list b4096 rc=$?"
# full capture of the dominant kernel: all 7 levels of the second step
$B32 > $OUT/plain_b32b.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:scan_tc -s 7 -c 7 -o $OUT/prof_scan_b32 -f $B32 > $OUT/ncu_full_b32.log 2>&1
echo "full b32 rc=$?"
$B4K > $OUT/plain_b4096b.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:scan_tc -s 7 -c 7 -o $OUT/prof_scan_b4096 -f $B4K > $OUT/ncu_full_b4096.log 2>&1
echo "full b4096 rc=$?"
ls -la $OUT
