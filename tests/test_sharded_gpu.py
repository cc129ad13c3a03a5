"""The merge kernel and the sharded search on real devices."""
import numpy as np
import pytest
import torch

from oracle import oracle
from tests.helpers import int_corpus

pytestmark = pytest.mark.gpu


def test_merge_kernel_matches_oracle():
    from kirag_b200.sharded import merge_topk_device

    rng = np.random.default_rng(0)
    for G, nq, k in ((2, 5, 10), (8, 33, 100), (4, 1, 1), (8, 3, 1024)):
        D_all = rng.integers(-5, 6, size=(G, nq, k)).astype(np.float32)  # many ties
        D_all = -np.sort(-D_all, axis=2)
        I_all = np.stack([np.stack([np.sort(rng.choice(10**6, k, replace=False)) + g * 10**6 for _ in range(nq)])
                          for g in range(G)]).astype(np.int64)
        I_all[-1, :, k - k // 3:] = -1  # padded tail of a short shard
        D_all[-1, :, k - k // 3:] = -3.4028234663852886e38
        Dm, Im = merge_topk_device(torch.from_numpy(D_all).cuda(), torch.from_numpy(I_all).cuda())
        Do, Io = oracle.merge_topk(D_all, I_all)
        # the oracle sorts by (score desc, id asc); ids inside a shard are ascending here, so both agree
        assert np.array_equal(Dm.cpu().numpy(), Do)
        assert np.array_equal(Im.cpu().numpy(), Io)


def test_host_pointer_merge_entry():
    import ctypes

    from kirag_b200 import _lib

    rng = np.random.default_rng(1)
    D_all = -np.sort(-rng.standard_normal((3, 4, 6)).astype(np.float32), axis=2)
    I_all = rng.permutation(3 * 4 * 6).reshape(3, 4, 6).astype(np.int64)
    D = np.empty((4, 6), dtype=np.float32)
    I = np.empty((4, 6), dtype=np.int64)
    p = lambda a: ctypes.c_void_p(a.ctypes.data)
    _lib.check(_lib.load().kirag_merge_topk(p(D_all), p(I_all), 3, 4, 6, p(D), p(I), 0, 0, None), "merge")
    Do, Io = oracle.merge_topk(D_all, I_all)
    assert np.array_equal(D, Do) and np.array_equal(I, Io)


def test_sharded_single_process_equals_unsharded():
    """G shards on one device, searched one after the other (no collective): merge == unsharded search."""
    from kirag_b200 import faiss_api
    from kirag_b200.sharded import merge_topk_device, shard_range

    rng = np.random.default_rng(2)
    n, d, k = 30000, 128, 50
    xb, xq = int_corpus(rng, n, d), int_corpus(rng, 9, d)
    full = faiss_api.IndexFlatIP(d)
    full.add(xb)
    D, I = full.search(xq, k)
    q = torch.from_numpy(xq).cuda()
    for G in (2, 4, 8):
        Ds, Is = [], []
        for r in range(G):
            lo, hi = shard_range(n, G, r)
            sh = faiss_api.IndexFlatIP(d)
            sh.add(xb[lo:hi])
            Dr, Ir = sh.search_device(q, k, id_offset=lo)
            Ds.append(Dr), Is.append(Ir)
        Dm, Im = merge_topk_device(torch.stack(Ds), torch.stack(Is))
        assert np.array_equal(Im.cpu().numpy(), I) and np.array_equal(Dm.cpu().numpy(), D)


def test_weighted_shards_single_process_equal_unsharded_and_probe_runs():
    """Speed-weighted row ranges (ShardedFlatIP(weights=...)): uneven shards, same global answer; the rank-speed probe
    (measure_rank_weights) runs on one GPU and returns [1.0] without a process group."""
    from kirag_b200 import faiss_api
    from kirag_b200.sharded import measure_rank_weights, merge_topk_device, shard_range

    rng = np.random.default_rng(3)
    n, d, k = 40000, 128, 20
    xb, xq = int_corpus(rng, n, d), int_corpus(rng, 5, d)
    full = faiss_api.IndexFlatIP(d)
    full.add(xb)
    D, I = full.search(xq, k)
    q = torch.from_numpy(xq).cuda()
    w = [1.3, 0.7, 1.05, 0.95]
    Ds, Is, sizes = [], [], []
    for r in range(4):
        lo, hi = shard_range(n, 4, r, w)
        sizes.append(hi - lo)
        sh = faiss_api.IndexFlatIP(d)
        sh.add(xb[lo:hi])
        Dr, Ir = sh.search_device(q, k, id_offset=lo)
        Ds.append(Dr), Is.append(Ir)
    assert sum(sizes) == n and sizes[0] > sizes[2] > sizes[3] > sizes[1]
    Dm, Im = merge_topk_device(torch.stack(Ds), torch.stack(Is))
    assert np.array_equal(Im.cpu().numpy(), I) and np.array_equal(Dm.cpu().numpy(), D)
    assert measure_rank_weights(128, 0, nq=64, k=10, rows=1 << 16, seconds=0.2) == [1.0]


def test_search_graph_without_peers_is_the_plain_search():
    """ShardedFlatIP.search_graph falls back to search() when there is no peer exchange (world size 1)."""
    from kirag_b200.sharded import ShardedFlatIP

    rng = np.random.default_rng(4)
    xb, xq = int_corpus(rng, 20000, 128), int_corpus(rng, 3, 128)
    sh = ShardedFlatIP(128, 20000, rank=0, world_size=1, device=0)
    sh.add_shard(torch.from_numpy(xb).cuda())
    q = torch.from_numpy(xq).cuda()
    D, I = sh.search(q, 10)
    Dg, Ig = sh.search_graph(q, 10)
    assert torch.equal(D, Dg) and torch.equal(I, Ig)
    Do, Io = oracle.flat_ip_search(xb, xq, 10)
    assert np.array_equal(I.cpu().numpy(), Io)
