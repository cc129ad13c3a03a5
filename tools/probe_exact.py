"""Time of the exact fp32 path (KIRAG_PATH_EXACT) per query group: python tools/probe_exact.py ROWS NQ1,NQ2,..."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import bench
from kirag_b200 import _lib, faiss_api
rows = int(sys.argv[1]); nqs = [int(x) for x in sys.argv[2].split(",")]
dev = torch.device("cuda", 0)
ix = faiss_api.IndexFlatIP(1024, device=0); ix.reserve(rows); bench.build_shard(ix, 0, rows, dev)
g = torch.Generator(device=dev); g.manual_seed(4321)
q_all = torch.nn.functional.normalize(torch.randn(max(nqs), 1024, generator=g, device=dev), dim=1)
for nq in nqs:
    q = q_all[:nq].contiguous()
    for _ in range(2): D, I = ix.search_device(q, 100, path=_lib.PATH_EXACT)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n = 5
    e0.record()
    for _ in range(n): D, I = ix.search_device(q, 100, path=_lib.PATH_EXACT)
    e1.record(); torch.cuda.synchronize()
    Da, Ia = ix.search_device(q, 100)
    print(f"rows={rows} nq={nq}: exact {e0.elapsed_time(e1)/n:.3f} ms  == auto: {bool(torch.equal(I, Ia) and torch.equal(D, Da))}", flush=True)
