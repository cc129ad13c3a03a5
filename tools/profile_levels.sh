#!/bin/bash
set -u
OUT=gpurun_out
mkdir -p $OUT
export PROBE_STEPS=1
CMD="python tools/probe.py 21000000 4096"
$CMD > $OUT/plain_levels.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:scan_tc -s 11 -c 3 -o $OUT/prof_levels -f $CMD > $OUT/ncu_levels.log 2>&1
echo "levels rc=$?"
