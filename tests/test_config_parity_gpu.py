"""Parity at the sizes the numbers are quoted on: BASELINE configs[2] (5.2M x 1024, batch 1024, top-100) and
configs[3] (21M x 1024, batches 1 / 32 / 4096, top-100) — the shapes of `index.search(q, top_docs)` at
/root/reference/retriever/index.py:47 that bench.py measures.

No host oracle finishes in seconds at these sizes, so the checks are the size-independent ones:
  * the certified tcgen05 filter path (KIRAG_PATH_AUTO) against the exact fp32 CUDA-core scan
    (KIRAG_PATH_EXACT, itself pinned to the oracle at small sizes in test_search_gpu.py) on >= 64 sampled
    queries per batch size: ids AND scores bit-equal;
  * self-retrieval of planted rows (query = a noisy copy of a known corpus row -> that row is rank 1);
  * every result row ordered by (score desc, id asc), ids in range, no duplicates;
  * the answer does not depend on the batch a query travels in (batch 1 == batch 32 == batch 4096 rows).
The corpus is bench.py's (chunk-seeded, generated on the device), so these are the benchmarked bytes.
"""
import gc

import numpy as np
import pytest
import torch

import bench
from kirag_b200 import _lib

pytestmark = pytest.mark.gpu

K = 100


def _free_device_memory():
    gc.collect()
    torch.cuda.empty_cache()


def _make_index(n_rows, reserve=True):
    from kirag_b200 import faiss_api

    _free_device_memory()
    free, _total = torch.cuda.mem_get_info(0)
    need = n_rows * 1024 * 6 + (12 << 30)
    if free < need:
        pytest.skip(f"needs {need >> 30} GiB of free HBM, {free >> 30} GiB available")
    ix = faiss_api.IndexFlatIP(1024, device=0)
    if reserve:
        ix.reserve(n_rows)
    bench.build_shard(ix, 0, n_rows, torch.device("cuda", 0))
    assert ix.ntotal == n_rows
    return ix


def _queries(ix, n_rows, batch, n_planted, seed=4321):
    """bench.py's query distribution; the first n_planted rows are noisy copies of known corpus rows."""
    dev = torch.device("cuda", 0)
    g = torch.Generator(device=dev)
    g.manual_seed(seed)
    q = torch.nn.functional.normalize(torch.randn(batch, 1024, generator=g, device=dev), dim=1)
    rows = np.linspace(17, n_rows - 19, n_planted).astype(np.int64)
    for j, r in enumerate(rows):
        x = torch.from_numpy(ix.reconstruct_n(int(r), 1)).to(dev)
        noise = torch.randn(1, 1024, generator=g, device=dev) / 32.0  # |noise| ~ 1: cosine to the row ~ 0.7
        q[j] = torch.nn.functional.normalize(x + 0.6 * noise, dim=1)[0]
    return q.contiguous(), rows


def _check_rows(D, I, n_rows):
    assert np.all((I >= 0) & (I < n_rows))
    assert np.all((D[:, :-1] > D[:, 1:]) | ((D[:, :-1] == D[:, 1:]) & (I[:, :-1] < I[:, 1:])))
    assert all(len(set(r.tolist())) == r.size for r in I)


def _auto_vs_exact(ix, q, sample, n_rows, what):
    D, I = ix.search_device(q, K)
    torch.cuda.synchronize()
    st = dict(ix.last_stats)
    assert st["n_overflow"] == 0, (what, st)
    D, I = D.cpu().numpy(), I.cpu().numpy()
    _check_rows(D, I, n_rows)
    idx = torch.as_tensor(sample, device=q.device)
    De, Ie = ix.search_device(q[idx].contiguous(), K, path=_lib.PATH_EXACT)
    torch.cuda.synchronize()
    assert ix.last_stats["n_exact"] == len(sample)
    De, Ie = De.cpu().numpy(), Ie.cpu().numpy()
    assert np.array_equal(I[sample], Ie), f"{what}: ids differ from the exact fp32 scan ({st})"
    assert np.array_equal(D[sample], De), f"{what}: scores differ from the exact fp32 scan ({st})"
    return D, I, st


def test_config2_5p2M_batch1024_top100():
    n, B = 5_200_000, 1024
    ix = _make_index(n)
    try:
        q, planted = _queries(ix, n, B, 16)
        sample = np.unique(np.concatenate([np.arange(16), np.linspace(16, B - 1, 64).astype(np.int64)]))
        D, I, st = _auto_vs_exact(ix, q, sample, n, "configs[2]")
        assert st["n_fast"] + st["n_rescan"] + st["n_exact"] + st["n_retry"] == B and st["n_fast"] >= B - 8, st
        assert np.array_equal(I[:16, 0], planted)
    finally:
        del ix
        _free_device_memory()


@pytest.fixture(scope="module")
def index_21m():
    # NO reserve: 21 add() calls of 2^20 rows, the way the reference's build path feeds an index
    # (faiss_index_corpus.py:42-46: one index_data per 1M-row file, no reserve hook).  The index ends at 129 GB of
    # the 180 GB, so growing by allocate-copy-free (two copies resident at once) could not get here.
    ix = _make_index(21_000_000, reserve=False)
    yield ix
    del ix
    _free_device_memory()


@pytest.mark.parametrize("batch", [32, 4096])
def test_config3_21M_top100(index_21m, batch):
    n = 21_000_000
    ix = index_21m
    n_planted = 8
    q, planted = _queries(ix, n, batch, n_planted)
    sample = np.arange(batch) if batch <= 64 else np.unique(
        np.concatenate([np.arange(n_planted), np.linspace(n_planted, batch - 1, 64).astype(np.int64)]))
    D, I, st = _auto_vs_exact(ix, q, sample, n, f"configs[3] batch {batch}")
    assert st["n_fast"] >= batch - max(1, batch // 128), st
    assert np.array_equal(I[:n_planted, 0], planted)
    # the answer to a query does not depend on the batch it travels in: batch 1, and (for 4096) batch 32
    for j in (0, n_planted, batch - 1):
        D1, I1 = ix.search_device(q[j:j + 1].contiguous(), K)
        torch.cuda.synchronize()
        assert np.array_equal(I1.cpu().numpy()[0], I[j]) and np.array_equal(D1.cpu().numpy()[0], D[j])
    if batch > 32:
        D32, I32 = ix.search_device(q[100:132].contiguous(), K)
        torch.cuda.synchronize()
        assert np.array_equal(I32.cpu().numpy(), I[100:132]) and np.array_equal(D32.cpu().numpy(), D[100:132])


def test_config3_21M_batch1_top100(index_21m):
    """KiRAG's real call shape (1-2 queries per retrieval, knowledge_graph/models.py:1645) at DPR scale."""
    n = 21_000_000
    ix = index_21m
    q, planted = _queries(ix, n, 8, 4, seed=99)
    for j in range(8):
        D, I = ix.search_device(q[j:j + 1].contiguous(), K)
        torch.cuda.synchronize()
        st = dict(ix.last_stats)
        De, Ie = ix.search_device(q[j:j + 1].contiguous(), K, path=_lib.PATH_EXACT)
        torch.cuda.synchronize()
        D, I = D.cpu().numpy(), I.cpu().numpy()
        _check_rows(D, I, n)
        assert np.array_equal(I, Ie.cpu().numpy()) and np.array_equal(D, De.cpu().numpy()), st
        if j < 4:
            assert I[0, 0] == planted[j]
    # host-pointer call (the faiss-shaped entry the reference uses) == device-pointer call
    Dh, Ih = ix.search(q.cpu().numpy(), K)
    Dd, Id = ix.search_device(q, K)
    torch.cuda.synchronize()
    assert np.array_equal(Ih, Id.cpu().numpy()) and np.array_equal(Dh, Dd.cpu().numpy())
