"""The reference's OWN Indexer (retriever/index.py), unmodified, on the oracle's faiss-shaped module.

Build-container only (needs /root/reference).  Pins the call-site glue the oracle is wrapped in:
fp32 cast, batches of index_batch_size, int64 id map, str ids, (ids, scores) tuples.
"""
import importlib
import os
import sys

import numpy as np
import pytest

from oracle import oracle
from tests.conftest import REFERENCE
from tests.helpers import int_corpus

pytestmark = pytest.mark.reference


@pytest.fixture()
def ref_index_module(have_reference):
    if not have_reference:
        pytest.skip("/root/reference not present")
    saved = {k: sys.modules.get(k) for k in ("faiss", "retriever", "retriever.index")}
    sys.modules["faiss"] = oracle.make_faiss_module()
    sys.path.insert(0, REFERENCE)
    sys.modules.pop("retriever.index", None)
    mod = importlib.import_module("retriever.index")
    yield mod
    sys.path.remove(REFERENCE)
    for k, v in saved.items():
        if v is None:
            sys.modules.pop(k, None)
        else:
            sys.modules[k] = v


def test_reference_indexer_runs_unmodified_on_the_oracle(ref_index_module, tmp_path):
    rng = np.random.default_rng(2)
    xb, xq = int_corpus(rng, 300, 32), int_corpus(rng, 5, 32)
    ids = [str(7 * i) for i in range(300)]
    ix = ref_index_module.Indexer(32, "inner_product")
    ix.index_data(ids[:100], xb[:100])
    ix.index_data(ids[100:], xb[100:])
    res = ix.search_knn(xq, 10, index_batch_size=2, verbose=False)
    D, I = oracle.flat_ip_search(xb, xq, 10)
    for r, (db_ids, scores) in enumerate(res):
        assert db_ids == [str(7 * i) for i in I[r]]
        assert np.array_equal(np.asarray(scores), D[r])
    ix.serialize(str(tmp_path))
    ix2 = ref_index_module.Indexer(32, "inner_product")
    ix2.deserialize_from(str(tmp_path))
    res2 = ix2.search_knn(xq, 10, verbose=False)
    assert [r[0] for r in res2] == [r[0] for r in res]


def test_mirror_and_reference_indexer_agree(ref_index_module, monkeypatch):
    from kirag_b200 import index as kindex

    monkeypatch.setitem(kindex.INDEX_TYPES, "inner_product", oracle.OracleIndexFlatIP)
    rng = np.random.default_rng(3)
    xb, xq = int_corpus(rng, 120, 16), int_corpus(rng, 2500, 16)  # > 2 batches of 1024
    ids = list(range(500, 620))
    a, b = ref_index_module.Indexer(16), kindex.Indexer(16)
    a.index_data(ids, xb), b.index_data(ids, xb)
    ra, rb = a.search_knn(xq, 10, verbose=False), b.search_knn(xq, 10, verbose=False)
    assert len(ra) == len(rb) == 2500
    for (ia, sa), (ib, sb) in zip(ra, rb):
        assert ia == ib and np.array_equal(sa, sb)
