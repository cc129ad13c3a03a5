#!/bin/bash
# DRAM bytes of every scan launch of the bench command at batch 4096 (single-pass metrics: no replay, no memory
# save/restore), under a few L2 policy settings.  usage: tools/profile_traffic.sh TAG "ENV1=a ENV2=b" ...
set -u
OUT=gpurun_out
T=$1; shift
B4K="python bench.py --steps 2 --warmup 1 --batch 4096 --sweep= --no-cpu-baseline --no-extras"
i=0
for SETTING in "$@"; do
  env $SETTING $B4K > $OUT/${T}_plain_$i.json 2> $OUT/${T}_plain_$i.err &&
  env $SETTING timeout 900 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none \
      -k regex:scan_tc -c 28 --csv --log-file $OUT/${T}_traffic_$i.csv $B4K > $OUT/${T}_ncu_$i.log 2>&1
  echo "setting $i [$SETTING] rc=$?"
  python - <<PY
import csv
from collections import defaultdict
rows=[r for r in csv.reader(open("$OUT/${T}_traffic_$i.csv",errors="replace")) if len(r)>14 and r[0].isdigit()]
per=defaultdict(dict)
for r in rows: per[int(r[0])][r[12]]=float(r[14].replace(",",""))
ids=sorted(per); steps=[ids[j:j+7] for j in range(0,len(ids),7)]; steps=[s for s in steps if len(s)==7]
for s in steps[-2:]:
    print("  step: read %.2f GB write %.2f GB time %.2f ms" % (sum(per[j]["dram__bytes_read.sum"] for j in s)/1e9, sum(per[j]["dram__bytes_write.sum"] for j in s)/1e9, sum(per[j]["gpu__time_duration.sum"] for j in s)/1e6))
PY
  i=$((i+1))
done
