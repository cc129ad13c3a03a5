"""ShardedIndexer under torchrun (one process per GPU): equals the single-GPU Indexer on every rank."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from kirag_b200 import Indexer, ShardedIndexer  # noqa: E402


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    rng = np.random.default_rng(5)
    n, d, nq, k = 120_000, 1024, 300, 20
    xb = rng.standard_normal((n, d)).astype(np.float32)
    xb /= np.linalg.norm(xb, axis=1, keepdims=True)
    xb[n // 3] = xb[2 * n // 3]  # exact tie across ranks: lower passage id first
    xq = rng.standard_normal((nq, d)).astype(np.float32)
    xq[0] = xb[n // 3]
    ids = [str(3 * i + 2) for i in range(n)]
    sh = ShardedIndexer(d, device=local)
    ref = Indexer(d, device=local)
    for a in range(0, n, 25_000):
        sh.index_data(ids[a:a + 25_000], xb[a:a + 25_000])
        ref.index_data(ids[a:a + 25_000], xb[a:a + 25_000])
    got = sh.search_knn(xq, k, index_batch_size=128, verbose=False)
    want = ref.search_knn(xq, k, index_batch_size=128, verbose=False)
    ok = len(got) == len(want) and all(g[0] == w[0] and np.array_equal(g[1], w[1]) for g, w in zip(got, want))
    flag = torch.tensor([1 if ok else 0], device=torch.device("cuda", local))
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    if rank == 0:
        print(f"local rows per rank ~{sh.index.ntotal}, exchange={sh._sh.exchange}, peer={'yes' if sh._sh.peer else 'no'}")
        print("sharded_indexer_check ok" if flag.item() else "sharded_indexer_check FAILED", flush=True)
    sh.close()
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0 if flag.item() else 1)


if __name__ == "__main__":
    main()
