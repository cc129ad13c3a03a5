"""Host-side logic that needs no GPU: shard ranges, embedding-file pairing, the Indexer glue, the faiss stand-in."""
import os
import pickle
import sys

import numpy as np
import pytest

from kirag_b200 import build_index, faiss_api
from kirag_b200 import index as kindex
from kirag_b200.sharded import shard_range
from oracle import oracle
from tests.helpers import int_corpus


def test_weighted_shard_ranges_partition_the_rows_in_proportion():
    """Speed-weighted shards (ShardedFlatIP(weights=...)): contiguous, disjoint, complete, tile-aligned inner
    boundaries, sizes proportional to the weights; equal weights ~ equal shards."""
    n = 21_000_000
    w = [1.07, 1.0, 1.01, 0.98, 0.96, 1.0, 0.95, 0.95]
    spans = [shard_range(n, 8, r, w) for r in range(8)]
    assert spans[0][0] == 0 and spans[-1][1] == n
    for (a0, a1), (b0, b1) in zip(spans, spans[1:]):
        assert a1 == b0 and a0 <= a1 and b0 % 128 == 0
    for (lo, hi), wi in zip(spans, w):
        assert abs((hi - lo) - n * wi / sum(w)) <= 256
    eq = [shard_range(n, 8, r, [1.0] * 8) for r in range(8)]
    assert all(abs((hi - lo) - n / 8) <= 256 for lo, hi in eq)
    # degenerate: fewer rows than ranks * tile
    tiny = [shard_range(300, 4, r, [1, 1, 1, 1]) for r in range(4)]
    assert tiny[0][0] == 0 and tiny[-1][1] == 300 and all(a1 == b0 for (_, a1), (b0, _) in zip(tiny, tiny[1:]))


def test_shard_ranges_partition_the_rows():
    for n in (0, 1, 7, 8, 9, 1000, 21_000_000):
        for G in (1, 2, 4, 8):
            spans = [shard_range(n, G, r) for r in range(G)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            for (a, b), (c, d) in zip(spans, spans[1:]):
                assert b == c and a <= b and c <= d


def test_pairing_is_exact_not_substring(tmp_path):
    # the reference pairs "…_0_999999" with "passage_id_list_1000000_1999999" by substring
    # (faiss_index_corpus.py:37-41); exact pairing must not
    names = ["corpus_embeddings_0_999999.pkl", "corpus_embeddings_1000000_1999999.pkl",
             "passage_id_list_1000000_1999999.pkl", "passage_id_list_0_999999.pkl"]
    for n in names:
        (tmp_path / n).write_bytes(b"x")
    pairs = build_index.pair_embedding_files(str(tmp_path))
    assert [(os.path.basename(a), os.path.basename(b)) for a, b in pairs] == [
        ("corpus_embeddings_0_999999.pkl", "passage_id_list_0_999999.pkl"),
        ("corpus_embeddings_1000000_1999999.pkl", "passage_id_list_1000000_1999999.pkl")]


def test_pairing_survives_dotted_directories(tmp_path):
    d = tmp_path / "e5.large.v2"
    d.mkdir()
    (d / "corpus_embeddings_0_9.pkl").write_bytes(b"x")
    (d / "passage_id_list_0_9.pkl").write_bytes(b"x")
    assert len(build_index.pair_embedding_files(str(d))) == 1


def test_faiss_stand_in_exposes_what_index_py_names():
    import kirag_b200.as_faiss as af

    mod = af.make_module()
    for name in ("IndexFlatIP", "IndexFlatL2", "IndexPQ", "METRIC_INNER_PRODUCT", "IO_FLAG_MMAP", "write_index",
                 "read_index"):
        assert hasattr(mod, name)
    with pytest.raises(NotImplementedError):
        mod.IndexFlatL2(8)
    with pytest.raises(NotImplementedError):
        mod.IndexPQ(8, 2, 8, mod.METRIC_INNER_PRODUCT)


class _OracleBackedIndex(oracle.OracleIndexFlatIP):
    """Test double: stands in for the CUDA index so the Indexer glue can run on CPU."""


def test_indexer_mirror_glue_on_a_test_double(tmp_path, monkeypatch):
    monkeypatch.setitem(kindex.INDEX_TYPES, "inner_product", _OracleBackedIndex)
    monkeypatch.setattr(faiss_api, "write_index", oracle._write_index)
    monkeypatch.setattr(faiss_api, "read_index", oracle._read_index)
    rng = np.random.default_rng(1)
    xb, xq = int_corpus(rng, 40, 16), int_corpus(rng, 3, 16)
    ids = [str(1000 + 3 * i) for i in range(40)]
    ix = kindex.Indexer(16)
    ix.index_data(ids[:25], xb[:25].astype(np.float64))  # any float dtype, like index.py:29
    ix.index_data(ids[25:], xb[25:])
    res = ix.search_knn(xq, 5, index_batch_size=2, verbose=False)
    D, I = oracle.flat_ip_search(xb, xq, 5)
    assert len(res) == 3
    for r, (db_ids, scores) in enumerate(res):
        assert db_ids == [str(1000 + 3 * i) for i in I[r]] and all(isinstance(s, str) for s in db_ids)
        assert np.array_equal(scores, D[r])
    ix.serialize(str(tmp_path))
    ix2 = kindex.Indexer(16)
    ix2.deserialize_from(str(tmp_path))
    assert ix2.index.ntotal == 40 and np.array_equal(ix2.index_id_to_db_id, ix.index_id_to_db_id)
    # k > ntotal: -1 padding maps to the LAST id, exactly as the reference's index_id_to_db_id[-1] does
    res = ix.search_knn(xq[:1], 45, verbose=False)
    assert res[0][0][-1] == ids[-1]


def test_ixfi_container_layout(tmp_path):
    idx = oracle.OracleIndexFlatIP(4)
    idx.add(np.arange(12, dtype=np.float32).reshape(3, 4))
    p = tmp_path / "index.faiss"
    oracle._write_index(idx, str(p))
    raw = p.read_bytes()
    assert raw[:4] == b"IxFI" and len(raw) == 45 + 3 * 4 * 4
    assert int.from_bytes(raw[4:8], "little") == 4 and int.from_bytes(raw[8:16], "little") == 3
    assert int.from_bytes(raw[16:24], "little") == 1 << 20 and raw[32] == 1
    assert int.from_bytes(raw[33:37], "little") == 0 and int.from_bytes(raw[37:45], "little") == 12
    assert np.array_equal(np.frombuffer(raw[45:], dtype="<f4"), np.arange(12, dtype=np.float32))


def test_embedding_shard_writer_matches_reference_file_format(tmp_path):
    """compute_corpus_embeddings.py:114-115 file pairs, written per rank without any collective, are
    consumed by the (exact-pairing) index builder's file discovery in ascending order."""
    import torch

    from kirag_b200.embed_writer import ContiguousShardSampler, EmbeddingShardWriter

    n, d, per_file, world = 2500, 8, 1000, 2
    emb = torch.arange(n * d, dtype=torch.float32).reshape(n, d)
    ids = [str(7 * i) for i in range(n)]
    for rank in range(world):
        sampler = ContiguousShardSampler(n, rank, world)
        w = EmbeddingShardWriter(str(tmp_path), d, sampler.lo, sampler.hi, ids, num_passage_per_index_file=per_file)
        rows = list(sampler)
        for i in range(0, len(rows), 96):  # batches that straddle file boundaries
            b = rows[i:i + 96]
            w.add(b, emb[b])
        w.close()
    pairs = build_index.pair_embedding_files(str(tmp_path))
    names = [os.path.basename(a) for a, _ in pairs]
    assert names == ["corpus_embeddings_0_999.pkl", "corpus_embeddings_1000_1249.pkl",
                     "corpus_embeddings_1250_1999.pkl", "corpus_embeddings_2000_2499.pkl"]
    got_e, got_i = [], []
    for a, b in pairs:
        t = pickle.load(open(a, "rb"))
        assert isinstance(t, torch.Tensor) and t.dtype == torch.float32 and not t.is_cuda
        got_e.append(t)
        got_i += pickle.load(open(b, "rb"))
    assert torch.equal(torch.cat(got_e), emb) and got_i == ids


def test_contiguous_sampler_covers_everything_once():
    from kirag_b200.embed_writer import ContiguousShardSampler

    for n in (0, 5, 1001):
        for world in (1, 3, 8):
            seen = [i for r in range(world) for i in ContiguousShardSampler(n, r, world)]
            assert seen == list(range(n))


# ------------------------------------------------------------- level schedule ---
def _schedule(n_rows, nq, k, d=1024):
    import ctypes

    from kirag_b200 import _lib

    lib = _lib.load()
    out = (ctypes.c_int64 * 64)()
    cap, kp = ctypes.c_int(), ctypes.c_int()
    n = lib.kirag_debug_level_schedule(n_rows, nq, k, d, out, 64, ctypes.byref(cap), ctypes.byref(kp))
    return n, [int(out[i]) for i in range(max(n, 0))], cap.value, kp.value


@pytest.mark.parametrize("n_rows", [1, 127, 128, 129, 4096, 20_000, 100_000, 430_000, 2_625_000, 5_200_000, 21_000_000])
@pytest.mark.parametrize("nq,k", [(1, 10), (2, 100), (32, 100), (64, 20), (128, 100), (129, 100), (1024, 10), (4096, 100),
                                  (16384, 512)])
def test_level_schedule_covers_the_corpus_with_bounded_growth(n_rows, nq, k):
    """The filter path's level schedule (csrc/api.cu::level_bounds), a pure host function: strictly
    increasing, ends at the whole corpus, level 0 is half the candidate buffer, later levels never
    grow the prefix by more than cap / (4 k') (expected survivors stay below a quarter of the buffer)."""
    n, hi, cap, kp = _schedule(n_rows, nq, k)
    assert n == len(hi) and n >= 1
    assert kp == min(max(4 * k, 32), 2048)
    assert cap == (32768 if nq <= 128 else 8192)
    assert hi[-1] == n_rows
    assert all(a < b for a, b in zip(hi, hi[1:]))
    assert hi[0] == min(n_rows, cap // 2)
    gmax = max(2, min(32, cap // (4 * kp)))
    tiles = [-(-h // 128) for h in hi]
    g1 = min(16 if (cap == 32768 and nq <= 8) else 4, gmax)  # level 1: 4x; 16x for calls of at most 8 queries
    for lvl, (a, b) in enumerate(zip(tiles, tiles[1:]), start=1):
        limit = g1 if lvl == 1 else gmax
        assert b <= a * limit + 1, (lvl, a, b, limit)  # +1 tile: ceil of the geometric step
    # and it does not use more levels than a greedy walk with the same limits would
    t, greedy = tiles[0], 1
    n_tiles = tiles[-1]
    if t < n_tiles:
        t, greedy = min(n_tiles, t * g1), 2
    while t < n_tiles:
        t, greedy = min(n_tiles, t * gmax), greedy + 1
    assert n <= greedy


def test_level_schedule_reference_points():
    # the shapes DESIGN.md quotes
    assert _schedule(21_000_000, 4096, 100)[0] == 7
    assert _schedule(21_000_000, 32, 100)[0] == 4
    assert _schedule(2_625_000, 4096, 100)[0] == 6
    assert _schedule(2_625_000, 32, 100)[0] == 4
    assert _schedule(2_625_000, 2, 100)[0] == 3    # one 8-GPU shard, KiRAG's call shape: 16k / 262k / 2.6M rows
    assert _schedule(5_200_000, 2, 10)[0] == 3
    assert _schedule(430_000, 64, 20)[0] <= 3
    assert _schedule(1000, 4, 10)[1] == [1000]


def test_level_schedule_ineligible_shapes():
    assert _schedule(1000, 4, 10, d=100)[0] == -1      # d not a multiple of 64: exact scan only
    # k > 512: the over-fetch is capped at k' = 2048, the filter path still takes it (k <= 2048)
    n_levels, bounds, cap, kprime = _schedule(1_000_000, 4, 1024)
    assert n_levels > 0 and kprime == 2048 and cap >= 4 * kprime and bounds[-1] == 1_000_000


def test_indexer_rejects_what_it_does_not_implement():
    with pytest.raises(NotImplementedError):
        kindex.Indexer(8, metric="l2")
    with pytest.raises(NotImplementedError):
        kindex.Indexer(8, n_subquantizers=4)


def test_passage_id_table_grows_amortised_and_accepts_strings():
    t = kindex._PassageIds()
    t.extend(["5", "6"])
    t.extend(np.arange(7, 3000))
    t.extend([])
    assert len(t) == 2995 and t.view()[:3].tolist() == [5, 6, 7] and t.view().dtype == np.int64


def test_triple_scorer_cache_reset_keeps_the_current_call_consistent():
    """when the bank overflows max_cached it is rebuilt from EVERYTHING the current call needs (texts that
    were cached before the reset included), not only from the texts that were new."""
    import torch

    from kirag_b200 import aligner

    calls = []

    def embed(texts):
        calls.append(list(texts))
        return torch.tensor([[float(x), 1.0] for x in texts])

    sc = aligner.TripleScorer(embed_queries=embed, embed_documents=embed, max_cached=5)
    # run the bookkeeping on CPU tensors: _ensure only moves tensors that are not CUDA, so patch .cuda() away
    orig_cuda = torch.Tensor.cuda
    torch.Tensor.cuda = lambda self, *a, **k: self
    try:
        a = sc._ensure(["1", "2", "3"])
        assert a[:, 0].tolist() == [1.0, 2.0, 3.0] and sc.n_embedded == 3
        b = sc._ensure(["3", "2"])  # cached, permuted: a gather
        assert b[:, 0].tolist() == [3.0, 2.0] and sc.n_embedded == 3
        c = sc._ensure(["2", "3", "4", "5", "6", "7"])  # 3 held + 4 new > 5: the bank starts over with all six
        assert c[:, 0].tolist() == [2.0, 3.0, 4.0, 5.0, 6.0, 7.0]
        assert calls[-1] == ["2", "3", "4", "5", "6", "7"] and len(sc._row_of) == 6
        d = sc._ensure(["4", "5"])  # contiguous run of the bank: a view, nothing embedded
        assert d[:, 0].tolist() == [4.0, 5.0] and sc.n_embedded == 9
    finally:
        torch.Tensor.cuda = orig_cuda
