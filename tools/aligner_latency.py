import sys, time, torch
sys.path.insert(0, '/root/repo')
from kirag_b200.scoring import topk_inner_product
dev = torch.device('cuda', 0)
g = torch.Generator(device=dev); g.manual_seed(5)
def ev(fn, n=200):
    for _ in range(20): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3
def wall(fn, n=200):
    for _ in range(20): fn()
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(n): fn(); 
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / n * 1e6
for C, T in ((1, 200), (2, 1000), (2, 5000), (8, 20000), (16, 20000), (256, 50000)):
    Tm = torch.nn.functional.normalize(torch.randn(T, 1024, generator=g, device=dev), dim=1)
    Q = torch.nn.functional.normalize(torch.randn(C, 1024, generator=g, device=dev), dim=1)
    k = min(20, T)
    ours = ev(lambda: topk_inner_product(Q, Tm, k)); ours_w = wall(lambda: topk_inner_product(Q, Tm, k))
    ref = ev(lambda: torch.topk(torch.matmul(Q, Tm.T), k=k, dim=1)); ref_w = wall(lambda: torch.topk(torch.matmul(Q, Tm.T), k=k, dim=1))
    Th, Qh = Tm.cpu(), Q.cpu()
    host = wall(lambda: topk_inner_product(Qh.numpy(), Th.numpy(), k), 20)
    refh = wall(lambda: torch.topk(torch.matmul(Qh, Th.T), k=k, dim=1), 20)
    print(f"C={C} T={T}: device ours {ours:.1f} us (wall {ours_w:.1f}) torch {ref:.1f} us (wall {ref_w:.1f}) | host buffers: ours {host:.0f} us, torch CPU (reference) {refh:.0f} us", flush=True)
