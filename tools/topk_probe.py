import os, sys, time, torch
sys.path.insert(0, '/root/repo')
from kirag_b200 import scoring
dev = torch.device('cuda', 0)
g = torch.Generator(device=dev); g.manual_seed(5)
T = torch.nn.functional.normalize(torch.randn(50_000, 1024, generator=g, device=dev), dim=1)
Q = torch.nn.functional.normalize(torch.randn(256, 1024, generator=g, device=dev), dim=1)
def ev(fn, n, w):
    for _ in range(w): fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(n):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); e1.synchronize(); ts.append(e0.elapsed_time(e1))
    ts.sort(); return ts[len(ts)//2]*1e3, ts[0]*1e3, ts[-1]*1e3
for env in ({}, {"KIRAG_NO_CENTER": "1"}):
    for k_, v in env.items(): os.environ[k_] = v
    print(env, "256x50000:", ev(lambda: scoring.topk_inner_product(Q, T, 20), 50, 5))
    Q16 = Q[:16].contiguous(); T20 = T[:20000].contiguous()
    print(env, "16x20000:", ev(lambda: scoring.topk_inner_product(Q16, T20, 20), 50, 5))
    print(env, "256x50000 again:", ev(lambda: scoring.topk_inner_product(Q, T, 20), 50, 5))
