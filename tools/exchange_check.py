"""Multi-GPU check of the row-sharded search (run under torchrun, one process per GPU):
peer-memory exchange kernel == NCCL all-gather + merge kernel == unsharded search on rank 0's GPU.
Also times both exchanges.  `python -m torch.distributed.run --nproc-per-node N tools/exchange_check.py`"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from kirag_b200 import faiss_api  # noqa: E402
from kirag_b200.sharded import ShardedFlatIP, shard_range  # noqa: E402


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dev = torch.device("cuda", local)
    n, d = int(os.environ.get("XCHK_ROWS", 200_000)), 1024
    g = torch.Generator(device=dev)
    g.manual_seed(99)  # same stream of numbers on every rank
    xb = torch.nn.functional.normalize(torch.randn(n, d, generator=g, device=dev), dim=1)
    xb[n // 3] = xb[n // 2 + 7]  # exact duplicates across shards: tie broken by lower id
    xq_all = torch.nn.functional.normalize(torch.randn(4096, d, generator=g, device=dev), dim=1)
    xq_all[0] = xb[n // 2 + 7]
    lo, hi = shard_range(n, world, rank)
    peer = ShardedFlatIP(d, n, rank=rank, world_size=world, device=local, exchange="peer", max_nq=4096, max_k=128)
    peer.add_shard(xb[lo:hi])
    nccl = ShardedFlatIP(d, n, rank=rank, world_size=world, device=local, exchange="nccl", local_index=peer.index)
    full = None
    if rank == 0:
        full = faiss_api.IndexFlatIP(d, device=local)
        full.add_device(xb)
    ok = True
    for nq, k in ((1, 10), (2, 100), (33, 20), (256, 100), (1024, 100), (4096, 100), (7, 128)):
        q = xq_all[:nq].contiguous()
        for rep in range(3):  # both parities of the double buffer
            Dp, Ip = peer.search(q, k)
        Dn, In = nccl.search(q, k)
        same = bool(torch.equal(Ip, In) and torch.equal(Dp, Dn))
        if nq <= 256:  # the whole chain replayed from one CUDA graph (twice: capture + replay, both parities)
            for rep in range(3):
                Dg, Ig = peer.search_graph(q, k)
                same = same and bool(torch.equal(Ig, Ip) and torch.equal(Dg, Dp))
        if rank == 0:
            Df, If = full.search_device(q, k)
            same = same and bool(torch.equal(Ip, If) and torch.equal(Dp, Df))
        flag = torch.tensor([1 if same else 0], device=dev)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        if rank == 0:
            print(f"nq={nq} k={k}: peer == nccl == unsharded on every rank: {bool(flag.item())}", flush=True)
        ok = ok and bool(flag.item())
        # timing of the exchange alone (per-shard results already there)
        D_loc, I_loc = peer.search_local(q, k)
        for name, fn in (("peer", lambda: peer.peer.merge(D_loc, I_loc)), ("nccl", lambda: nccl_exchange(nccl, D_loc, I_loc))):
            for _ in range(3):
                fn()
            dist.barrier()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(20):
                fn()
            e1.record()
            torch.cuda.synchronize()
            t = torch.tensor([e0.elapsed_time(e1) / 20], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            if rank == 0:
                print(f"    exchange {name}: {t.item() * 1e3:8.1f} us (max over {world} ranks)", flush=True)
    # Clustered rows on ONE shard only: that rank's certificates fail, the others' pass.  The flags travel with
    # the exchange, every rank sees the same OR word, the flagged rank re-answers and all ranks exchange again.
    centers = torch.nn.functional.normalize(torch.randn(8, d, generator=g, device=dev), dim=1)
    pick = torch.randint(0, 8, (n,), generator=g, device=dev)
    xc = xb.clone()
    c_lo, c_hi = shard_range(n, world, world - 1)
    xc[c_lo:c_hi] = torch.nn.functional.normalize(
        centers[pick[c_lo:c_hi]] + 0.0008 * torch.randn(c_hi - c_lo, d, generator=g, device=dev), dim=1)
    qc = torch.nn.functional.normalize(centers[:4] + 0.002 * torch.randn(4, d, generator=g, device=dev), dim=1)
    clustered = ShardedFlatIP(d, n, rank=rank, world_size=world, device=local, exchange="peer", max_nq=64, max_k=128)
    clustered.add_shard(xc[lo:hi])
    Dc, Ic = clustered.search(qc, 10)
    st = dict(clustered.index.last_stats)
    De, Ie = clustered.search(qc, 10, path=1)  # exact fp32 scan on every shard + the same exchange
    same = bool(torch.equal(Ic, Ie) and torch.equal(Dc, De))
    flagged = torch.tensor([st.get("n_cert_fail", 0) + st.get("n_overflow", 0)], device=dev)
    allf = [torch.zeros_like(flagged) for _ in range(world)]
    dist.all_gather(allf, flagged)
    flag = torch.tensor([1 if same else 0], device=dev)
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    if rank == 0:
        per_rank = [int(x.item()) for x in allf]
        print(f"clustered shard {world - 1}: flagged queries per rank {per_rank}; async search + flag-carrying exchange == "
              f"exact on every rank: {bool(flag.item())}", flush=True)
        ok = ok and per_rank[-1] > 0
    ok = ok and bool(flag.item())
    clustered.close()
    peer.close()
    dist.barrier()
    if rank == 0:
        print("exchange_check ok" if ok else "exchange_check FAILED", flush=True)
    dist.destroy_process_group()
    sys.exit(0 if ok else 1)


def nccl_exchange(sh, D_loc, I_loc):
    nq, k = D_loc.shape
    packed = torch.empty((2, nq, k), dtype=torch.int64, device=I_loc.device)
    packed[0] = D_loc.contiguous().view(torch.int32).to(torch.int64)
    packed[1] = I_loc
    gathered = torch.empty((sh.world_size, 2, nq, k), dtype=torch.int64, device=I_loc.device)
    dist.all_gather_into_tensor(gathered.view(-1, k), packed.view(-1, k), group=sh.group)
    D_all = gathered[:, 0].to(torch.int32).view(torch.float32).contiguous()
    I_all = gathered[:, 1].contiguous()
    return sh.merge_fn(D_all, I_all)


if __name__ == "__main__":
    main()
