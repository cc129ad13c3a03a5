"""Digest ncu artefacts (gpurun_out/*.ncu-rep, launch-list CSVs) into the small text/JSON files kept under profiles/.

    python tools/ncu_digest.py launches <launches.csv> <first-kernel-regex>      -> per-kernel share of one step
    python tools/ncu_digest.py full <report.ncu-rep>                             -> one line per captured launch
"""
import csv
import re
import subprocess
import sys
from collections import OrderedDict

UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}
TIME = {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3, "nsecond": 1e-6, "usecond": 1e-3, "msecond": 1.0, "second": 1e3}


def short(name: str) -> str:
    name = re.sub(r"\(.*", "", name)
    name = name.replace("void ", "").replace("kirag::", "")
    return name[:60]


def launches(path: str, first: str):
    """Launch list of `bench.py`: take the LAST complete search (from a launch matching `first` up to, but
    excluding, the next one) and print every kernel's device time and share."""
    rows = [r for r in csv.reader(open(path, errors="replace")) if len(r) > 14 and r[0].isdigit()]
    # a list captured with several metrics has one line per (launch, metric): keep the duration lines
    if any(r[12] != "gpu__time_duration.sum" for r in rows):
        rows = [r for r in rows if r[12] == "gpu__time_duration.sum"]
    names = [r[4] for r in rows]
    starts = [i for i, n in enumerate(names) if re.search(first, n)]
    assert len(starts) >= 2, "need at least two searches in the launch list"
    a, b = starts[-2], starts[-1]
    step = rows[a:b]
    agg = OrderedDict()
    total = 0.0
    for r in step:
        ms = float(r[14]) * TIME.get(r[13], 1e-6)
        k = short(r[4])
        n, t = agg.get(k, (0, 0.0))
        agg[k] = (n + 1, t + ms)
        total += ms
    print(f"# one search = launches {rows[a][0]}..{rows[b - 1][0]} of {path}: {len(step)} launches, {total:.4f} ms of kernel time")
    print(f"# (gpu__time_duration.sum per launch under ncu: cold caches, serialised — compare SHARES, not absolutes)")
    for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"{k:62s} x{n:<3d} {t:10.4f} ms  {100.0 * t / total:6.2f} %")


WANT = OrderedDict([
    ("gpu__time_duration.sum", "ms"),
    ("dram__bytes_read.sum", "GB"),
    ("dram__bytes_write.sum", "GB"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "%"),
    ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "%"),
    ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "%"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "%"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "%"),
    ("launch__registers_per_thread", ""),
    ("launch__shared_mem_per_block_dynamic", "KB"),
    ("sm__cycles_elapsed.avg.per_second", "GHz"),
])


def full(path: str):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    cols = {h: i for i, h in enumerate(hdr)}
    print("# " + path)
    print("# kernel | grid | " + " | ".join(f"{k.split('.')[0]}[{u}]" for k, u in WANT.items()))
    for r in rows[2:]:
        vals = []
        for k, u in WANT.items():
            if k not in cols:
                vals.append("-")
                continue
            v, un = r[cols[k]], units[cols[k]].split("/")[0]
            try:
                x = float(v)
            except ValueError:
                vals.append(v)
                continue
            if u == "GB":
                x = x * UNIT.get(un, 1.0) / 1e9
            elif u == "ms":
                x = x * TIME.get(un, 1.0)
            elif u == "KB":
                x = x * UNIT.get(un, 1.0) / 1e3
            vals.append(f"{x:.4g}")
        print(f"{short(r[cols['Kernel Name']]):44s} | {r[cols['Grid Size']]:14s} | " + " | ".join(vals))


if __name__ == "__main__":
    if sys.argv[1] == "launches":
        launches(sys.argv[2], sys.argv[3])
    else:
        full(sys.argv[2])
