"""world_size-2 gloo test of the sharded-search plumbing (shard ranges, global ids, gather layout, merge call).

The CUDA local search and merge kernel are replaced by oracle-backed test doubles here; the real
kernels are covered by tests/test_sharded_gpu.py.
"""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import oracle
from tests.helpers import int_corpus


class _LocalOracleIndex:
    def __init__(self, d):
        self.d = d
        self.x = np.empty((0, d), dtype=np.float32)

    @property
    def ntotal(self):
        return self.x.shape[0]

    def add(self, x):
        self.x = np.concatenate([self.x, np.asarray(x, dtype=np.float32)])

    def search_device(self, q, k, id_offset=0):
        D, I = oracle.flat_ip_search(self.x, q.numpy(), k, id_offset=id_offset)
        return torch.from_numpy(D), torch.from_numpy(I)


def _merge(D_all, I_all):
    D, I = oracle.merge_topk(D_all.numpy(), I_all.numpy())
    return torch.from_numpy(D), torch.from_numpy(I)


def _worker(rank, world, port, n, d, nq, k, ret, weights=None):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from kirag_b200.sharded import ShardedFlatIP

        rng = np.random.default_rng(11)
        xb, xq = int_corpus(rng, n, d), int_corpus(rng, nq, d)
        sh = ShardedFlatIP(d, n, local_index=_LocalOracleIndex(d), merge_fn=_merge, weights=weights)
        sh.add_shard(xb[sh.lo:sh.hi])
        D, I = sh.search(torch.from_numpy(xq), k)
        D1, I1 = oracle.flat_ip_search(xb, xq, k)
        ret[rank] = bool(np.array_equal(I.numpy(), I1) and np.array_equal(D.numpy(), D1))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("n,k", [(101, 7), (3, 5)])  # second case: shards smaller than k -> padded partial results
def test_sharded_search_world_size_2(n, k):
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ret = mp.Manager().dict()
    mp.spawn(_worker, args=(2, port, n, 16, 4, k, ret), nprocs=2, join=True)
    assert ret[0] and ret[1]


def test_sharded_search_world_size_2_with_speed_weighted_shards():
    """Unequal row ranges (ShardedFlatIP(weights=...), what bench.py uses at N > 1): same global result."""
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ret = mp.Manager().dict()
    mp.spawn(_worker, args=(2, port, 1000, 16, 4, 9, ret, [1.3, 0.7]), nprocs=2, join=True)
    assert ret[0] and ret[1]


# ---------------------------------------------------------------- ShardedIndexer ---
def _indexer_worker(rank, world, port, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from kirag_b200.index import ShardedIndexer

        rng = np.random.default_rng(21)
        d, k = 16, 6
        chunks = [(list(range(a, b)), int_corpus(rng, b - a, d)) for a, b in ((0, 40), (40, 45), (45, 120), (120, 121), (121, 200))]
        xq = int_corpus(rng, 7, d)
        ix = ShardedIndexer(d, local_index=_LocalOracleIndex(d), merge_fn=_merge)
        for ids, emb in chunks:
            ix.index_data([str(i) for i in ids], emb)  # ids arrive as strings, like the reference's passage ids
        got = ix.search_knn(xq, k, index_batch_size=4)
        xb = np.concatenate([e for _, e in chunks])
        D1, I1 = oracle.flat_ip_search(xb, xq, k)
        ok = len(got) == 7 and ix.ntotal_global == 200
        for (ids, scores), d_ref, i_ref in zip(got, D1, I1):
            ok = ok and ids == [str(i) for i in i_ref] and np.array_equal(scores, d_ref)
        # every chunk lives on exactly one rank
        ok = ok and ix.index.ntotal == sum(len(ids) for c, (ids, _) in enumerate(chunks) if c % world == rank)
        ret[rank] = bool(ok)
    finally:
        dist.destroy_process_group()


def test_sharded_indexer_world_size_2_matches_single_index():
    """SPMD `Indexer` over two ranks (chunks dealt round-robin, passage ids mapped before the merge) returns what
    one flat index over all rows returns: same ids (as str), same scores, same (score desc, id asc) order."""
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ret = mp.Manager().dict()
    mp.spawn(_indexer_worker, args=(2, port, ret), nprocs=2, join=True)
    assert ret[0] and ret[1]
