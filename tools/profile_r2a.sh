#!/bin/bash
# Round 2, first GPU call: evidence for the mid-batch scan (128 < B <= 512) and the small-batch latency
# chain BEFORE changing either.  Every ncu pass follows the identical plain
# command having exited 0; numbers printed under ncu are never bench values.
set -u
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,power.limit --format=csv > $OUT/r2a_smi.log 2>&1
# plain timing + per-launch rates + phase timeline at the full 21M rows (mid batches) and on an 8-GPU shard
PROBE_STEPS=5 python tools/probe.py 21000000 128,256,512,1024 100 > $OUT/r2a_probe_21m.log 2>&1
echo "probe 21m rc=$?"
PROBE_STEPS=10 python tools/probe.py 2625000 1,32,256,4096 100 > $OUT/r2a_probe_shard.log 2>&1
echo "probe shard rc=$?"
# ncu --set full of the LAST (largest) scan level at B=256 and B=512 on a 2.625M-row corpus (one 8-GPU shard) (ncu saves and
# restores device memory around every replay pass: the 129 GB corpus would take ~8 min per capture)
for B in 256 512; do
  CMD="python tools/probe.py 2625000 $B 100"
  PROBE_STEPS=1 $CMD > $OUT/r2a_plain_b$B.log 2>&1 &&
  PROBE_STEPS=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:"scan_tc" -s 28 -c 2 \
      -o $OUT/r2a_scan_b$B -f $CMD > $OUT/r2a_ncu_b$B.log 2>&1
  echo "ncu full b$B rc=$?"
done
ls -la $OUT | grep r2a
