#!/bin/bash
# Same-box A/B of two builds of the library (KIRAG_B200_LIB): alternating runs of tools/probe_knobs.py.
# usage: tools/ab_libs.sh LIB_A LIB_B ROWS BATCHES [ROUNDS]
A=$1; B=$2; ROWS=$3; BATCHES=$4; ROUNDS=${5:-2}
for r in $(seq 1 $ROUNDS); do
  for L in "$A" "$B"; do
    echo "== round $r lib $L"
    KIRAG_B200_LIB=$L PROBE_STEPS=${PROBE_STEPS:-20} PROBE_REPS=1 python tools/probe_knobs.py $ROWS $BATCHES "KIRAG_SCAN_MULTI=1" 2>&1 | grep "^rep" | cut -c30-90
  done
done
