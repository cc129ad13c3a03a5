#!/bin/bash
# Round 2, after the threshold-prefetch change in the streamed 2-CTA scan: bench line, launch list + DRAM traffic of the
# bench command, ncu --set full of the last scan level at B = 128 / 256 / 512 / 4096 on a 2.625M-row corpus (one 8-GPU
# shard; ncu saves and restores device memory around every replay pass, the 129 GB corpus would take minutes per capture).
# Every ncu pass follows the identical plain command having exited 0; numbers printed under ncu are never bench values.
set -u
OUT=gpurun_out
T=${1:-r2m}
mkdir -p $OUT
python bench.py > $OUT/${T}_bench.json 2> $OUT/${T}_bench.err; echo "bench rc=$?"
B4K="python bench.py --steps 2 --warmup 1 --batch 4096 --sweep= --no-cpu-baseline --no-extras"
$B4K > $OUT/${T}_plain_b4096.json 2> $OUT/${T}_plain_b4096.err &&
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv \
    --log-file $OUT/${T}_launches_b4096.csv $B4K > $OUT/${T}_ncu_launch.log 2>&1
echo "launch list rc=$?"
# DRAM bytes of every scan launch of the same command (single-pass metrics: no replay, no memory save/restore)
timeout 900 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none \
    -k regex:scan_tc -c 60 --csv --log-file $OUT/${T}_traffic_b4096.csv $B4K > $OUT/${T}_ncu_traffic.log 2>&1
echo "traffic rc=$?"
for B in 128 256 512 4096; do
  CMD="python tools/probe.py 2625000 $B 100"
  PROBE_STEPS=1 $CMD > $OUT/${T}_plain_b$B.log 2>&1 &&
  PROBE_STEPS=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:"scan_tc" -s 28 -c 2 \
      -o $OUT/${T}_scan_b$B -f $CMD > $OUT/${T}_ncu_b$B.log 2>&1
  echo "ncu full b$B rc=$?"
done
ls -la $OUT | grep ${T}
