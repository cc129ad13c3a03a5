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
