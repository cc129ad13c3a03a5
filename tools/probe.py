"""Quick perf probe on one GPU: python tools/probe.py ROWS B1,B2,... [K] — prints ms/step and scan-kernel rates."""
import ctypes
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import bench  # noqa: E402
from kirag_b200 import _lib, faiss_api  # noqa: E402

rows = int(sys.argv[1]) if len(sys.argv) > 1 else 4_000_000
batches = [int(b) for b in (sys.argv[2] if len(sys.argv) > 2 else "32,4096").split(",")]
k = int(sys.argv[3]) if len(sys.argv) > 3 else 100
steps = int(os.environ.get("PROBE_STEPS", 5))
dev = torch.device("cuda", 0)
lib = _lib.load()
ix = faiss_api.IndexFlatIP(1024, device=0)
ix.reserve(rows)
t0 = time.time()
bench.build_shard(ix, 0, rows, dev)
print(f"built {rows} rows in {time.time() - t0:.1f}s", flush=True)
g = torch.Generator(device=dev)
g.manual_seed(4321)
q_all = torch.nn.functional.normalize(torch.randn(max(batches), 1024, generator=g, device=dev), dim=1)
for B in batches:
    q = q_all[:B].contiguous()
    for _ in range(3):
        ix.search_device(q, k)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        ix.search_device(q, k)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    lib.kirag_profile_enable(1)
    for _ in range(steps):
        ix.search_device(q, k)
    torch.cuda.synchronize()
    per_ms = (ctypes.c_double * 64)()
    per_rows = (ctypes.c_double * 64)()
    lib.kirag_profile_read_launches(per_ms, per_rows, 64)
    tl_tags = (ctypes.c_int * 512)()
    tl_ms = (ctypes.c_double * 512)()
    n_tl = lib.kirag_profile_read_timeline(tl_tags, tl_ms, 512)
    sm, ln, rw = ctypes.c_double(), ctypes.c_int64(), ctypes.c_double()
    lib.kirag_profile_read(ctypes.byref(sm), ctypes.byref(ln), ctypes.byref(rw))
    lib.kirag_profile_enable(0)
    scan = sm.value / steps
    gbs = rows * 2048 / (scan * 1e-3) / 1e9
    tf = 2.0 * B * rows * 1024 / (scan * 1e-3) / 1e12
    print(f"B={B:6d} k={k} step {ms:9.3f} ms  qps {B / ms * 1e3:10.1f}  scan {scan:9.3f} ms ({scan / ms:.3f} of step, "
          f"{ln.value // steps} launches)  {gbs:8.1f} GB/s  {tf:8.1f} TFLOP/s  stats {ix.last_stats}", flush=True)
    nl = ln.value // steps
    last = [(per_ms[i], per_rows[i]) for i in range(ln.value - nl, ln.value)] if ln.value <= 64 else []
    print("      last step per launch: " + "  ".join(
        f"[{int(r)} rows {m:.3f} ms {r * 2048 / m / 1e6:.0f} GB/s {2.0 * B * r * 1024 / m / 1e9:.0f} TF]" for m, r in last), flush=True)
    per_step = n_tl // steps if steps else 0
    if per_step:
        base = (steps - 1) * per_step
        t0 = tl_ms[base]
        print("      timeline (last step, ms since start): " + " ".join(
            f"{tl_tags[base + i]}:{tl_ms[base + i] - t0:.3f}" for i in range(per_step)), flush=True)
