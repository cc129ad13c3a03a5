#!/bin/bash
# ncu evidence for the FINAL round-1 kernels (run under gpurun, 1 GPU).  Every ncu pass follows the
# identical plain command having exited 0.  Numbers printed under ncu are never bench values.
# Launch lists use the bench command itself (21M rows); the --set full captures use one 8-GPU shard
# (2,625,000 rows) because ncu saves/restores device memory around every replay pass and the full
# corpus (129 GB) makes each capture take ~8 minutes.
set -u
OUT=gpurun_out
mkdir -p $OUT
B32="python bench.py --steps 2 --warmup 1 --batch 32 --sweep= --no-cpu-baseline"
B4K="python bench.py --steps 2 --warmup 1 --batch 4096 --sweep= --no-cpu-baseline"
$B32 > $OUT/plain3_b32.log 2>&1 &&
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 150 -c 400 --csv --log-file $OUT/launches3_b32.csv $B32 > $OUT/ncu3_launches_b32.log 2>&1
echo "launch list b32 rc=$?"
$B4K > $OUT/plain3_b4096.log 2>&1 &&
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 150 -c 400 --csv --log-file $OUT/launches3_b4096.csv $B4K > $OUT/ncu3_launches_b4096.log 2>&1
echo "launch list b4096 rc=$?"
S32="python bench.py --rows 2625000 --steps 2 --warmup 1 --batch 32 --sweep= --no-cpu-baseline"
S4K="python bench.py --rows 2625000 --steps 2 --warmup 1 --batch 4096 --sweep= --no-cpu-baseline"
$S32 > $OUT/plain3_s32.log 2>&1 &&
timeout 400 ncu --set full --clock-control none --import-source on -k regex:"scan_tc|compact_topm|rescore|final_kernel" -s 10 -c 10 -o $OUT/prof3_shard_b32 -f $S32 > $OUT/ncu3_full_s32.log 2>&1
echo "full shard b32 rc=$?"
$S4K > $OUT/plain3_s4096.log 2>&1 &&
timeout 500 ncu --set full --clock-control none --import-source on -k regex:"scan_tc|compact_topm|rescore|final_kernel" -s 14 -c 14 -o $OUT/prof3_shard_b4096 -f $S4K > $OUT/ncu3_full_s4096.log 2>&1
echo "full shard b4096 rc=$?"
ls -la $OUT | grep -E "3_|prof3" 
