"""Per-kernel histogram of the SASS opcodes that prove which hardware paths the shipped library uses
(tcgen05 MMA: UTCHMMA; TMEM: LDTM / UTCATOM...; bulk-copy engine: UBLKCP; tensor-core barriers: UTCBAR; mbarrier:
SYNCS; cluster: UCGABAR / MEMBAR...).  `python tools/sass_histogram.py > profiles/rNN_sass_opcodes.txt` (CPU only)."""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "kirag_b200", "libkirag_b200.so")
WANT = re.compile(r"\b(UTCHMMA|UTCQMMA|UTCOMMA|UTCBAR|UTCCP|UTCATOMSWS|LDTM|STTM|UBLKCP|UBLKPF|UTMALDG|UTMASTG|SYNCS|UCGABAR|"
                  r"CCTL|ACQBULK|ATOMG|ATOMS|REDG|RED|LDG|STG|LDS|STS|LDSM|HMMA|FFMA|BAR|MEMBAR|ERRBAR|ELECT|NANOSLEEP|SHFL|VOTE)"
                  r"(\.[A-Z0-9_.]+)?")


def main():
    out = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    kernels = collections.OrderedDict()
    cur = None
    for line in out.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip() or m.group(1)
            name = re.sub(r"\(.*", "", name).replace("kirag::", "").replace("void ", "")
            cur = kernels.setdefault(name, collections.Counter())
            continue
        if cur is None or "/*" not in line:
            continue
        body = line.split("*/", 1)[1] if "*/" in line else line
        for m in WANT.finditer(body):
            op = m.group(1) + (m.group(2) or "")
            cur[op.rstrip(".")] += 1
            cur["_instructions"] += 0
        if re.search(r"^\s+/\*[0-9a-f]{4}\*/", line):
            cur["_instructions"] += 1
    print(f"# SASS opcode histogram of {os.path.relpath(LIB, ROOT)} ({os.path.getsize(LIB)} bytes), cuobjdump -sass, sm_100a")
    print("# per kernel: total instructions, then the counts of the opcodes of interest")
    for name, c in kernels.items():
        total = c.pop("_instructions", 0)
        keys = sorted(c, key=lambda k_: (-c[k_], k_))
        tensor = [k_ for k_ in keys if k_.startswith(("UTC", "LDTM", "STTM", "UBLK", "UTMA", "SYNCS", "UCGA", "ACQBULK", "ELECT"))]
        rest = [k_ for k_ in keys if k_ not in tensor]
        print(f"\n{name}  [{total} instructions]")
        if tensor:
            print("    tcgen05 / TMEM / bulk-copy / mbarrier / cluster: " + "  ".join(f"{k_} x{c[k_]}" for k_ in tensor))
        print("    other: " + "  ".join(f"{k_} x{c[k_]}" for k_ in rest[:18]))


if __name__ == "__main__":
    sys.exit(main())
