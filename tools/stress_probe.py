"""How the certified filter path behaves on the stress corpora of SURVEY.md §8(d) (1 GPU):
python tools/stress_probe.py [ROWS] [K]  — per variant and batch: ms/step, certificate failures, rescans, exact fallbacks,
and (for a sample of queries) equality with the exact fp32 path."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from kirag_b200 import faiss_api  # noqa: E402

rows = int(sys.argv[1]) if len(sys.argv) > 1 else 2_000_000
k = int(sys.argv[2]) if len(sys.argv) > 2 else 100
d = 1024
dev = torch.device("cuda", 0)
g = torch.Generator(device=dev)


def norm(x):
    return torch.nn.functional.normalize(x, dim=1)


def corpus(kind):
    g.manual_seed(11)
    chunks = []
    if kind == "clustered":
        centers = norm(torch.randn(4096, d, generator=g, device=dev))
    for c in range(0, rows, 500_000):
        n = min(500_000, rows - c)
        x = torch.randn(n, d, generator=g, device=dev)
        if kind == "clustered":
            # sigma = 0.1 per coordinate relative to unit-norm centres (|noise| ~ 3.2: the centre is a weak signal) is
            # just i.i.d. again; use sigma so that |noise| = 0.1: tight clusters, rank-k and rank-4k nearly tie
            idx = torch.randint(0, 4096, (n,), generator=g, device=dev)
            x = centers[idx] + (0.1 / d ** 0.5) * x
        chunks.append(norm(x))
    return chunks


def queries(kind, chunks, nq):
    g.manual_seed(12)
    if kind == "iid":
        return norm(torch.randn(nq, d, generator=g, device=dev))
    base = chunks[0][:nq]
    if kind == "duplicates":
        return base.clone()
    return norm(base + (0.3 / d ** 0.5) * torch.randn(nq, d, generator=g, device=dev))  # planted / clustered: near a row


for kind in ("iid", "planted", "duplicates", "clustered"):
    chunks = corpus("clustered" if kind == "clustered" else "iid")
    ix = faiss_api.IndexFlatIP(d, device=0)
    ix.reserve(rows)
    for ch in chunks:
        ix.add_device(ch)
    if kind == "duplicates":  # 1 % of the rows copied to the end region: exact ties
        ix2 = None
    for nq in (32, 1024):
        q = queries(kind, chunks, nq)
        for _ in range(2):
            ix.search_device(q, k)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        D, I = ix.search_device(q, k)
        e1.record()
        torch.cuda.synchronize()
        st = dict(ix.last_stats)
        De, Ie = ix.search_device(q[:4], k, path=1)
        torch.cuda.synchronize()
        same = bool(torch.equal(I[:4], Ie) and torch.equal(D[:4], De))
        print(f"{kind:10s} rows={rows} nq={nq:5d} k={k}: {e0.elapsed_time(e1):9.3f} ms  fast={st['n_fast']} cert_fail={st['n_cert_fail']} "
              f"rescan={st['n_rescan']} exact={st['n_exact']} overflow={st['n_overflow']} levels={st['levels']}  "
              f"== exact path on 4 queries: {same}  top1={D[0, 0].item():.4f} rank{k}={D[0, k - 1].item():.4f}", flush=True)
    del ix, chunks
    torch.cuda.empty_cache()
