"""§8(f) "next" rows on the GPU: device-resident retriever (n2) and cached aligner scoring (n3), plus
search shapes at the edges of the level / buffer heuristics."""
import numpy as np
import pytest
import torch

from oracle import oracle
from tests.conftest import unit_rows
from tests.helpers import assert_topk_parity, int_corpus

pytestmark = pytest.mark.gpu


class _FakeEncoder(torch.nn.Module):
    """Stands in for BaseRetriever: embeds a 'text' (here: a row number as string) by table lookup."""

    def __init__(self, table):
        super().__init__()
        self.table = torch.from_numpy(table).cuda()
        self.device = self.table.device
        self.calls = 0

    def query(self, inputs):
        self.calls += 1
        return self.table[inputs["rows"]]

    doc = query


class _FakeCollator:
    def encode_query(self, texts, max_length=None, **kw):
        return {"rows": torch.tensor([int(t) for t in texts], dtype=torch.int64)}

    encode_doc = encode_query


class _Corpus:
    def __init__(self, n):
        self.docs = {str(1000 + i): {"id": str(1000 + i), "text": f"passage {i}"} for i in range(n)}

    def get_document(self, docid):
        return self.docs[docid]


def test_device_dense_retriever_matches_host_path():
    from kirag_b200 import Indexer
    from kirag_b200.retriever import DeviceDenseRetriever

    rng = np.random.default_rng(0)
    xb, qtab = unit_rows(rng, 30000, 256), unit_rows(rng, 50, 256)
    ix = Indexer(256)
    ix.index_data([str(1000 + i) for i in range(30000)], xb)
    enc = _FakeEncoder(qtab)
    r = DeviceDenseRetriever(enc, _FakeCollator(), indexer=ix, corpus=_Corpus(30000), batch_size=8)
    queries = [str(i) for i in range(20)]
    res = r(queries, topk=10)
    assert enc.calls == 3  # batches of 8
    Do, Io = oracle.flat_ip_search(xb, qtab[:20], 10, accum="f64")
    assert len(res) == 20 and all(len(x) == 10 for x in res)
    for qi in range(20):
        assert [d["id"] for d in res[qi]] == [str(1000 + i) for i in Io[qi]]
        np.testing.assert_allclose([d["score"] for d in res[qi]], Do[qi], rtol=1e-5, atol=1e-6)
        assert res[qi][0]["text"].startswith("passage")
    one = r("3", topk=5)  # str query -> single result list (retrievers.py:288-289)
    assert [d["id"] for d in one] == [str(1000 + i) for i in Io[3][:5]]
    # host path of the same interface (what the reference's DenseRetriever does) gives the same ids
    knn = ix.search_knn(qtab[:20], 10, verbose=False)
    assert [k[0] for k in knn] == [[d["id"] for d in res[qi]] for qi in range(20)]
    # no corpus -> {"id", "score"} dicts (retrievers.py:271-272)
    r2 = DeviceDenseRetriever(enc, _FakeCollator(), indexer=ix, corpus=None)
    assert set(r2(["0"], topk=3)[0][0].keys()) == {"id", "score"}


def test_triple_scorer_caches_and_matches_reference_expression():
    from kirag_b200.aligner import TripleScorer

    rng = np.random.default_rng(1)
    ttab, qtab = unit_rows(rng, 400, 128), unit_rows(rng, 4, 128)
    enc = _FakeEncoder(np.concatenate([ttab, qtab]))
    col = _FakeCollator()
    embed = lambda texts: enc.query(col.encode_query(texts))
    sc = TripleScorer(embed_queries=embed, embed_documents=embed)
    turn1 = [str(i) for i in range(0, 150)]
    turn2 = [str(i) for i in range(100, 400)]  # 50 already seen
    for texts in (turn1, turn2, turn2):
        idx, scores = sc.filter_candidate_triples([str(400), str(401)], texts, 20)
        rows = [int(t) for t in texts]
        ref_s, ref_i = oracle.topk_matmul_torch(torch.from_numpy(qtab[:2]), torch.from_numpy(ttab[rows]), 20)
        np.testing.assert_allclose(np.asarray(scores), ref_s.numpy(), rtol=1e-5, atol=1e-6)
        assert [set(r) for r in idx] == [set(r.tolist()) for r in ref_i]
    assert sc.n_embedded == 400  # every triple went through the encoder exactly once
    idx, scores = sc.filter_candidate_triples([str(402)], [str(5), str(6), str(7)], 20)  # fewer triples than k
    assert len(idx[0]) == 3


@pytest.mark.parametrize("n,d,nq,k", [(60000, 64, 5, 32), (60000, 64, 5, 33), (60000, 128, 3, 128), (50000, 64, 2, 512),
                                      (9000, 64, 17000, 3), (70000, 768, 40, 10), (40000, 192, 129, 7),
                                      (40000, 2048, 4, 10), (300000, 64, 2, 1)])
def test_search_shapes_at_heuristic_edges(n, d, nq, k):
    from kirag_b200 import faiss_api

    rng = np.random.default_rng(n + k + nq)
    xb, xq = int_corpus(rng, n, d), int_corpus(rng, nq, d)
    ix = faiss_api.IndexFlatIP(d)
    ix.add(xb)
    D, I, st = ix.search_ex(xq, k)
    sample = slice(0, min(nq, 64))
    assert_topk_parity(D[sample], I[sample], xb, xq[sample], k, exact=True, what=str(st))
    if nq > 64:
        De, Ie = oracle.flat_ip_search_blas(xb, xq, k, use_torch=True)
        assert np.array_equal(I, Ie) and np.array_equal(D, De)


def test_unit_vectors_large_k_and_odd_dim():
    from kirag_b200 import faiss_api

    rng = np.random.default_rng(5)
    xb, xq = unit_rows(rng, 120000, 320), unit_rows(rng, 6, 320)
    ix = faiss_api.IndexFlatIP(320)
    ix.add(xb)
    for k in (1, 100, 400, 1000):
        D, I, st = ix.search_ex(xq, k)
        assert_topk_parity(D, I, xb, xq, k, what=f"k={k} {st}")
    xb2, xq2 = unit_rows(rng, 20000, 100), unit_rows(rng, 3, 100)  # d % 64 != 0: exact path only
    ix2 = faiss_api.IndexFlatIP(100)
    ix2.add(xb2)
    D, I, st = ix2.search_ex(xq2, 10)
    assert st["n_exact"] == 3
    assert_topk_parity(D, I, xb2, xq2, 10, what="d=100")
