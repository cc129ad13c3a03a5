#!/bin/bash
set -u
OUT=gpurun_out
mkdir -p $OUT
export PROBE_STEPS=1
CMD="python tools/probe.py 4000000 4096"
KIRAG_SCAN_PAIR=1 $CMD > $OUT/plain_pair.log 2>&1 &&
KIRAG_SCAN_PAIR=1 ncu --set full --clock-control none --import-source on -k regex:scan_tc -s 23 -c 1 -o $OUT/prof_pair -f $CMD > $OUT/ncu_pair.log 2>&1
echo "pair rc=$?"
KIRAG_SCAN_PAIR=0 $CMD > $OUT/plain_single.log 2>&1 &&
KIRAG_SCAN_PAIR=0 ncu --set full --clock-control none --import-source on -k regex:scan_tc -s 23 -c 1 -o $OUT/prof_single -f $CMD > $OUT/ncu_single.log 2>&1
echo "single rc=$?"
tail -2 $OUT/plain_pair.log $OUT/plain_single.log
