import os, sys, torch
sys.path.insert(0, '/root/repo')
from kirag_b200 import pooling
dev = torch.device('cuda', 0)
g = torch.Generator(device=dev); g.manual_seed(777)
def timeit(fn, n=200):
    for _ in range(20): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3
for B in (1024, 512, 256, 128, 64, 8):
    for dt in (torch.float32, torch.bfloat16):
        h = torch.randn(B, 512, 1024, generator=g, device=dev).to(dt)
        for ragged in (False, True):
            if ragged:
                lens = torch.randint(1, 513, (B,), generator=torch.Generator().manual_seed(778)).to(dev)
                m = (torch.arange(512, device=dev)[None, :] < lens[:, None]).to(torch.int64)
            else:
                m = torch.ones(B, 512, dtype=torch.int64, device=dev)
            nbytes = int(m.sum().item()) * 1024 * h.element_size() + B * 512 * 8 + B * 1024 * 4
            res = []
            for cl in ('', '1', '2', '4', '8'):
                if cl: os.environ['KIRAG_POOL_CL'] = cl
                else: os.environ.pop('KIRAG_POOL_CL', None)
                us = timeit(lambda: pooling.e5_embed(h, m))
                res.append(f"cl={cl or 'auto'} {us:.1f}us {nbytes/us/1e3:.0f}GB/s")
            print(B, str(dt)[6:], 'ragged' if ragged else 'full', ' | '.join(res), flush=True)
