"""Where the time of a small-batch sharded search goes (run under torchrun, one process per GPU):
   python -m torch.distributed.run --nproc-per-node N tools/probe_sharded_latency.py [ROWS_PER_RANK] [B1,B2,...]
For every batch size, per step (max over ranks of the CUDA-event time, wall clock for the host view):
   local      IndexFlatIP.search_device_async alone, one sync at the very end (GPU-side floor of the local chain)
   chain      search_device_async + exchange kernel, one sync at the very end (GPU-side floor of a sharded search)
   search     ShardedFlatIP.search: the public call (one stream sync + flag check per call)
   graph      ShardedFlatIP.search_graph: the same chain replayed from a CUDA graph (one launch + one sync per call)
"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

import bench  # noqa: E402
from kirag_b200.sharded import ShardedFlatIP  # noqa: E402


def timed(fn, steps, dev, sync_each):
    for _ in range(5):
        fn()
    torch.cuda.synchronize(dev)
    dist.barrier()
    torch.cuda.synchronize(dev)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record()
    for _ in range(steps):
        fn()
        if sync_each:
            torch.cuda.current_stream(dev).synchronize()
    e1.record()
    torch.cuda.synchronize(dev)
    wall = (time.perf_counter() - t0) * 1e3 / steps
    t = torch.tensor([e0.elapsed_time(e1) / steps, wall], dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t[0]), float(t[1])


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dev = torch.device("cuda", local)
    per = int(sys.argv[1]) if len(sys.argv) > 1 else 2_625_000
    batches = [int(b) for b in (sys.argv[2] if len(sys.argv) > 2 else "1,32").split(",")]
    steps = int(os.environ.get("PROBE_STEPS", 50))
    k = 100
    n = per * world
    sh = ShardedFlatIP(1024, n, rank=rank, world_size=world, device=local, max_nq=4096, max_k=128)
    bench.build_shard(sh.index, sh.lo, sh.hi, dev)
    g = torch.Generator(device=dev)
    g.manual_seed(4321)
    q_all = torch.nn.functional.normalize(torch.randn(max(batches), 1024, generator=g, device=dev), dim=1)
    for B in batches:
        q = q_all[:B].contiguous()
        D_ref, I_ref = sh.search(q, k)

        def local_only():
            sh.index.search_device_async(q, k, id_offset=sh.lo)

        def chain():
            D_loc, I_loc = sh.index.search_device_async(q, k, id_offset=sh.lo)
            sh.peer.merge(D_loc, I_loc, sh.index.pending_flags_ptr())

        res = {}
        res["local"] = timed(local_only, steps, dev, False)
        sh.index.finish()
        res["chain"] = timed(chain, steps, dev, False)
        sh.index.finish()
        res["search"] = timed(lambda: sh.search(q, k), steps, dev, False)
        if hasattr(sh, "search_graph"):
            Dg, Ig = sh.search_graph(q, k)
            same = bool(torch.equal(Dg, D_ref) and torch.equal(Ig, I_ref))
            res["graph"] = timed(lambda: sh.search_graph(q, k), steps, dev, False)
        else:
            same = None
        if rank == 0:
            print(f"N={world} rows/rank={per} B={B}: " + "  ".join(
                f"{name} {ev:.3f} ms (host {wall:.3f})" for name, (ev, wall) in res.items()) +
                (f"  graph == search: {same}" if same is not None else ""), flush=True)
    sh.close()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
