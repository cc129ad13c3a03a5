#!/bin/bash
# Round 2, final kernels: the HBM regime (batch 32 -> scan_tc_pair_kernel<64,resident>): launch list of the bench command at
# batch 32, ncu --set full of its four scan levels at 21M rows (the resident kernel's replay passes re-read 43 GB each, the
# capture takes a few minutes), and the final bench line.  Every ncu pass follows the identical plain command having exited 0.
set -u
OUT=gpurun_out
T=${1:-r3n}
mkdir -p $OUT
B32="python bench.py --steps 2 --warmup 1 --batch 32 --sweep= --no-cpu-baseline --no-extras"
$B32 > $OUT/${T}_plain_b32.json 2> $OUT/${T}_plain_b32.err &&
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 300 --csv \
    --log-file $OUT/${T}_launches_b32.csv $B32 > $OUT/${T}_ncu_launch_b32.log 2>&1
echo "launch list b32 rc=$?"
CMD="python tools/probe.py 2625000 32 100"
PROBE_STEPS=1 $CMD > $OUT/${T}_plain_shard_b32.log 2>&1 &&
PROBE_STEPS=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:"scan_tc" -s 16 -c 4 \
    -o $OUT/${T}_scan_b32 -f $CMD > $OUT/${T}_ncu_b32.log 2>&1
echo "ncu full b32 rc=$?"
python bench.py > $OUT/${T}_bench.json 2> $OUT/${T}_bench.err; echo "bench rc=$?"
ls -la $OUT | grep ${T}
