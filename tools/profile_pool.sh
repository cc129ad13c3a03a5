#!/bin/bash
# ncu --set full of the pooling epilogue at [256,512,1024] fp32 and bf16 (the plain command first).
set -u
OUT=gpurun_out
T=${1:-r4e}
CMD="python tools/bench_pool.py"
$CMD > $OUT/${T}_pool_plain.log 2>&1 && timeout 600 ncu --set full --clock-control none --import-source on -k regex:pool_normalize_kernel -c 12 \
    -o $OUT/${T}_pool -f $CMD > $OUT/${T}_pool_ncu.log 2>&1
echo "ncu pool rc=$?"; tail -5 $OUT/${T}_pool_plain.log
