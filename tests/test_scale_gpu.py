"""Larger shapes: BASELINE configs[0]/[1] against the oracle, and size-independent properties at a few million rows."""
import numpy as np
import pytest

from oracle import oracle
from tests.conftest import unit_rows
from tests.helpers import assert_topk_parity

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def fa():
    from kirag_b200 import faiss_api

    assert faiss_api.get_num_gpus() >= 1
    return faiss_api


def test_config0_100k_x_1024_1k_queries_top10(fa):
    """BASELINE configs[0] (the CPU-FAISS config): ids against the blocked-sgemm oracle."""
    rng = np.random.default_rng(100)
    xb, xq = unit_rows(rng, 100_000, 1024), unit_rows(rng, 1000, 1024)
    ix = fa.IndexFlatIP(1024)
    ix.add(xb)
    D, I, st = ix.search_ex(xq, 10)
    assert st["n_fast"] >= 990, st
    Do, Io = oracle.flat_ip_search_blas(xb, xq, 10, use_torch=True)
    np.testing.assert_allclose(D, Do, rtol=1e-5, atol=1e-6)
    mism = int((I != Io).sum())
    assert mism <= 5, f"{mism} id differences vs the oracle"  # only fp32 near-tie swaps are tolerated
    assert_topk_parity(D[:50], I[:50], xb, xq[:50], 10, what="config0")


def test_config1_430k_x_1024_batch64_top20(fa):
    rng = np.random.default_rng(101)
    xb, xq = unit_rows(rng, 430_000, 1024), unit_rows(rng, 64, 1024)
    ix = fa.IndexFlatIP(1024)
    ix.reserve(430_000)
    for a in range(0, 430_000, 100_000):
        ix.add(xb[a:a + 100_000])
    D, I, st = ix.search_ex(xq, 20)
    assert st["n_fast"] == 64, st
    assert_topk_parity(D, I, xb, xq, 20, what=f"config1 {st}")


def test_properties_at_2M_rows(fa):
    """Size-independent properties on a device-generated 2M x 1024 corpus (no host oracle at this size):
    self-retrieval, row order, shard-merge == unsharded, fast == exact for a sample of queries."""
    import torch
    from kirag_b200.sharded import merge_topk_device

    n, d = 2_000_000, 1024
    ix = fa.IndexFlatIP(d)
    ix.reserve(n)
    g = torch.Generator(device="cuda")
    for c in range(0, n, 500_000):
        g.manual_seed(1234 + c)
        x = torch.randn(500_000, d, generator=g, device="cuda")
        x = torch.nn.functional.normalize(x, dim=1)
        ix.add_device(x)
        if c == 500_000:
            planted = x[1000:1016].clone()  # global rows 501000..501015
    q = torch.nn.functional.normalize(planted + 0.05 * torch.randn(16, d, device="cuda"), dim=1)
    D, I = ix.search_device(q, 100)
    torch.cuda.synchronize()
    st = ix.last_stats
    assert st["n_overflow"] == 0 and st["n_fast"] == 16, st
    D, I = D.cpu().numpy(), I.cpu().numpy()
    assert np.array_equal(I[:, 0], np.arange(501000, 501016))
    assert np.all((D[:, :-1] > D[:, 1:]) | ((D[:, :-1] == D[:, 1:]) & (I[:, :-1] < I[:, 1:])))
    De, Ie = ix.search_device(q[:4], 100, path=1)
    torch.cuda.synchronize()
    assert np.array_equal(Ie.cpu().numpy(), I[:4]) and np.array_equal(De.cpu().numpy(), D[:4])
    # two "shards" searched separately and merged == the unsharded answer (exactness of the sharded design)
    half = fa.IndexFlatIP(d)
    half.reserve(1_000_000)
    half.add_device(torch.from_numpy(ix.reconstruct_n(0, 1_000_000)).cuda())
    Da, Ia = half.search_device(q, 100, id_offset=0)
    other = fa.IndexFlatIP(d)
    other.reserve(1_000_000)
    other.add_device(torch.from_numpy(ix.reconstruct_n(1_000_000, 1_000_000)).cuda())
    Db, Ib = other.search_device(q, 100, id_offset=1_000_000)
    Dm, Im = merge_topk_device(torch.stack([Da, Db]), torch.stack([Ia, Ib]))
    torch.cuda.synchronize()
    assert np.array_equal(Im.cpu().numpy(), I) and np.array_equal(Dm.cpu().numpy(), D)
