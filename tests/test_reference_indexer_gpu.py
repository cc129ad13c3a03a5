"""The Indexer interface end to end on the GPU: our mirror, and (when a copy of the reference file is
present) the reference's own retriever/index.py running unmodified on kirag_b200.as_faiss."""
import os

import numpy as np
import pytest

from oracle import oracle
from tests.conftest import unit_rows

pytestmark = pytest.mark.gpu


def test_indexer_mirror_on_gpu(tmp_path):
    from kirag_b200 import Indexer

    rng = np.random.default_rng(0)
    xb, xq = unit_rows(rng, 20000, 1024), unit_rows(rng, 2100, 1024)
    ids = [str(10 * i + 3) for i in range(20000)]
    ix = Indexer(1024, "inner_product")
    ix.index_data(ids[:7000], xb[:7000])
    ix.index_data(ids[7000:], xb[7000:])
    res = ix.search_knn(xq, 10, index_batch_size=1024, verbose=False)
    Do, Io = oracle.flat_ip_search_blas(xb, xq, 10, use_torch=True)
    assert len(res) == 2100
    bad = 0
    for r, (db_ids, scores) in enumerate(res):
        assert len(db_ids) == 10 and isinstance(db_ids[0], str) and scores.dtype == np.float32
        np.testing.assert_allclose(scores, Do[r], rtol=1e-5, atol=1e-6)
        bad += db_ids != [str(10 * i + 3) for i in Io[r]]
    assert bad <= 3  # fp32 near-tie swaps only
    ix.serialize(str(tmp_path))
    assert sorted(os.listdir(tmp_path)) == ["index.faiss", "index_meta.faiss"]
    ix2 = Indexer(1024, "inner_product")
    ix2.deserialize_from(str(tmp_path))
    res2 = ix2.search_knn(xq[:64], 10, verbose=False)
    assert [r[0] for r in res2] == [r[0] for r in res[:64]]


def test_build_index_from_saved_embedding_shards(tmp_path):
    """faiss_index_corpus.py:27-52 flow: pickled torch tensors + id lists -> index.faiss + index_meta.faiss."""
    import pickle

    import torch

    from kirag_b200 import Indexer
    from kirag_b200.build_index import build_faiss_index

    rng = np.random.default_rng(1)
    xb = unit_rows(rng, 2500, 1024)
    for s, e in ((0, 999), (1000, 1999), (2000, 2499)):
        pickle.dump(torch.from_numpy(xb[s:e + 1]), open(tmp_path / f"corpus_embeddings_{s}_{e}.pkl", "wb"))
        pickle.dump([str(i) for i in range(s, e + 1)], open(tmp_path / f"passage_id_list_{s}_{e}.pkl", "wb"))
    build_faiss_index(index_folder=str(tmp_path), embedding_size=1024)
    assert sorted(os.listdir(tmp_path)) == ["index.faiss", "index_meta.faiss"]  # inputs deleted like the reference
    ix = Indexer(1024)
    ix.deserialize_from(str(tmp_path))
    xq = unit_rows(rng, 3, 1024)
    res = ix.search_knn(xq, 5, verbose=False)
    Do, Io = oracle.flat_ip_search(xb, xq, 5, accum="f64")
    assert [r[0] for r in res] == [[str(i) for i in row] for row in Io]


def test_reference_indexer_unmodified_on_the_b200_library(tmp_path):
    """retriever/index.py from the sha256-verified snapshot baseline/_ref (the only copy of the reference that
    exists on the GPU box), byte-for-byte unmodified, with `faiss` = kirag_b200.as_faiss."""
    from tests import refenv

    root = refenv.snapshot_root()
    if root is None:
        pytest.skip("baseline/_ref snapshot missing or stale: run __graft_entry__.build() in the build container")
    import kirag_b200.as_faiss as af
    from kirag_b200 import faiss_api

    with refenv.reference_modules(root, af.make_module()) as ref:
        rng = np.random.default_rng(2)
        xb, xq = unit_rows(rng, 50000, 1024), unit_rows(rng, 1100, 1024)  # two faiss calls of <= 1024 queries
        ix = ref.index.Indexer(1024, "inner_product")
        assert isinstance(ix.index, faiss_api.IndexFlatIP)
        ix.index_data([str(3 * i + 1) for i in range(20000)], xb[:20000])
        ix.index_data([str(3 * i + 1) for i in range(20000, 50000)], xb[20000:])
        res = ix.search_knn(xq, 10, verbose=False)
        Do, Io = oracle.flat_ip_search_blas(xb, xq, 10, use_torch=True)
        assert len(res) == 1100 and isinstance(res[0][0][0], str)
        bad = sum(r[0] != [str(3 * i + 1) for i in row] for r, row in zip(res, Io))
        assert bad <= 2  # fp32 near-tie swaps only
        np.testing.assert_allclose(np.stack([r[1] for r in res]), Do, rtol=1e-5, atol=1e-6)
        ix.serialize(str(tmp_path))
        ix2 = ref.index.Indexer(1024, "inner_product")
        ix2.deserialize_from(str(tmp_path))
        res2 = ix2.search_knn(xq[:40], 10, verbose=False)
        assert [r[0] for r in res2] == [r[0] for r in res[:40]]
        # k > ntotal: FAISS padding (-1) indexes the LAST id in the reference's own mapping (index.py:49)
        tiny = ref.index.Indexer(1024, "inner_product")
        tiny.index_data(["7", "8", "9"], xb[:3])
        (ids, scores), = tiny.search_knn(xq[:1], 5, verbose=False)
        assert ids[3:] == ["9", "9"] and np.all(scores[3:] == np.float32(-3.4028234663852886e38))


def test_sharded_indexer_single_rank_equals_indexer(tmp_path):
    """ShardedIndexer with one rank (no exchange): device-side row -> passage-id mapping, padding, serialise round
    trip; the two-rank plumbing is covered by tests/test_sharded_gloo.py and the exchange by tests/test_exchange_gpu.py."""
    from kirag_b200 import Indexer
    from kirag_b200.index import ShardedIndexer

    rng = np.random.default_rng(3)
    xb, xq = unit_rows(rng, 30000, 256), unit_rows(rng, 70, 256)
    ids = [str(7 * i + 1) for i in range(30000)]
    ref = Indexer(256, "inner_product")
    sh = ShardedIndexer(256, rank=0, world_size=1)
    for a, b in ((0, 12000), (12000, 12001), (12001, 30000)):
        ref.index_data(ids[a:b], xb[a:b])
        sh.index_data(ids[a:b], xb[a:b])
    want = ref.search_knn(xq, 20, index_batch_size=32, verbose=False)
    got = sh.search_knn(xq, 20, index_batch_size=32, verbose=False)
    assert len(got) == len(want) == 70
    for (gi, gs), (wi, ws) in zip(got, want):
        assert gi == wi and np.array_equal(gs, ws)
    sh.serialize(str(tmp_path))
    sh2 = ShardedIndexer(256, rank=0, world_size=1)
    sh2.deserialize_from(str(tmp_path))
    assert [r[0] for r in sh2.search_knn(xq[:8], 20, verbose=False)] == [r[0] for r in want[:8]]
    # more results requested than rows: padding ids stay "-1"
    tiny = ShardedIndexer(256, rank=0, world_size=1)
    tiny.index_data(ids[:3], xb[:3])
    (pi, ps), = tiny.search_knn(xq[:1], 5, verbose=False)
    assert pi[3:] == ["-1", "-1"] and np.all(ps[3:] == np.float32(-3.4028234663852886e38))
