#!/bin/bash
# ncu evidence, second pass of round 1 (run under gpurun, 1 GPU).  Each ncu pass is preceded by the
# identical plain command, which must exit 0 first.  Numbers printed under ncu are never bench values.
set -u
OUT=gpurun_out
mkdir -p $OUT
B32="python bench.py --steps 2 --warmup 1 --batch 32 --sweep= --no-cpu-baseline"
B4K="python bench.py --steps 2 --warmup 1 --batch 4096 --sweep= --no-cpu-baseline"
# launch lists (every launch of the bench command with its device time)
$B32 > $OUT/plain2_b32.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -s 150 -c 400 --csv --log-file $OUT/launches2_b32.csv $B32 > $OUT/ncu2_launches_b32.log 2>&1
echo "launch list b32 rc=$?"
$B4K > $OUT/plain2_b4096.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -s 150 -c 400 --csv --log-file $OUT/launches2_b4096.csv $B4K > $OUT/ncu2_launches_b4096.log 2>&1
echo "launch list b4096 rc=$?"
# full capture of the dominant kernel: the 7 levels of the second step
ncu --set full --clock-control none --import-source on -k regex:scan_tc -s 7 -c 7 -o $OUT/prof2_scan_b32 -f $B32 > $OUT/ncu2_full_b32.log 2>&1
echo "full b32 rc=$?"
ncu --set full --clock-control none --import-source on -k regex:scan_tc -s 7 -c 7 -o $OUT/prof2_scan_b4096 -f $B4K > $OUT/ncu2_full_b4096.log 2>&1
echo "full b4096 rc=$?"
# the small kernels of one step at batch 4096: compaction, rescoring, final
ncu --set full --clock-control none --import-source on -k regex:"compact_topm|rescore|final_kernel" -s 9 -c 9 -o $OUT/prof2_small_b4096 -f $B4K > $OUT/ncu2_small_b4096.log 2>&1
echo "small b4096 rc=$?"
# one 8-GPU shard (2.625M rows) at batch 4096: the early, survivor-dense levels with source attribution
SH="python tools/probe.py 2625000 4096"
PROBE_STEPS=1 $SH > $OUT/plain2_shard.log 2>&1 &&
PROBE_STEPS=1 ncu --set full --clock-control none --import-source on -k regex:scan_tc -s 24 -c 4 -o $OUT/prof2_shard_levels -f $SH > $OUT/ncu2_shard.log 2>&1
echo "shard rc=$?"
ls -la $OUT | tail -20
