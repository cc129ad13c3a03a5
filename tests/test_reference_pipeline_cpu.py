"""CPU twin of tests/test_reference_pipeline_gpu.py: the reference's own DenseRetriever / build_faiss_index /
Indexer, unmodified, on the ORACLE's faiss-shaped module, with the same tiny encoder, tokenizer and corpus.

It pins the host-side flow the GPU test relies on (file naming, id mapping, str ids, result dicts) and keeps
the shared harness (tests/refenv.py) exercised where there is no GPU.  `cal_doc_embeddings` itself hard-codes
`cuda:0` (compute_corpus_embeddings.py:52), so here the embedding files come from this repo's per-rank writer
driven by the reference's encoder; the GPU test runs the reference's producer itself.
"""
import argparse
import os

import numpy as np
import pytest
import torch

from oracle import oracle
from tests import refenv
from tests.helpers import assert_topk_parity

pytestmark = pytest.mark.reference

N_DOCS, DIM, TOPK = 240, 64, 10


@pytest.fixture()
def ref():
    root = refenv.any_reference_root()
    if root is None:
        pytest.skip("neither baseline/_ref nor /root/reference is present")
    with refenv.reference_modules(root, oracle.make_faiss_module()) as ns:
        yield ns


def test_snapshot_is_taken_and_matches_the_reference():
    """tools/snapshot_reference.py: baseline/_ref holds byte-identical copies (sha256 manifest)."""
    if not os.path.isdir(os.path.join(refenv.REFERENCE, "retriever")):
        pytest.skip("/root/reference not present")
    import sys

    sys.path.insert(0, os.path.join(refenv.ROOT, "tools"))
    try:
        import snapshot_reference as snap
    finally:
        sys.path.pop(0)
    dest = snap.snapshot()
    assert dest and snap.verify(dest)
    for rel in ("retriever/index.py", "retriever/retrievers.py", "retriever/encoders.py",
                "compute_corpus_embeddings.py", "faiss_index_corpus.py"):
        assert open(os.path.join(dest, rel), "rb").read() == open(os.path.join(refenv.REFERENCE, rel), "rb").read()
    # never part of the history
    ignored = open(os.path.join(refenv.ROOT, ".gitignore")).read()
    assert "baseline/_ref/" in ignored


def test_reference_retriever_flow_on_the_oracle(ref, tmp_path):
    with refenv.digit_free_dir() as scratch:
        _flow(ref, tmp_path, scratch)


def _flow(ref, tmp_path, scratch):
    from kirag_b200.embed_writer import ContiguousShardSampler, EmbeddingShardWriter

    docs = refenv.synthetic_docs(N_DOCS)
    tok = refenv.tiny_tokenizer()
    refenv.save_tiny_encoder(ref, str(tmp_path / "model"), hidden=DIM)
    collator = ref.collators.E5Collator(tokenizer=tok, query_maxlength=32, doc_maxlength=64)
    corpus = refenv.make_corpus_class(ref, docs)(title_prefix="title: ", passage_prefix="text: ")
    model = ref.retrievers.InBatchRetriever("E5Retriever", str(tmp_path / "model"), local_rank=-1, temperature=0.01)
    model.eval()
    folder = os.path.join(scratch, "idx")
    for rank in range(2):
        sampler = ContiguousShardSampler(len(corpus), rank, 2)
        loader = torch.utils.data.DataLoader(corpus, batch_size=8, sampler=sampler)
        w = EmbeddingShardWriter(str(folder), DIM, sampler.lo, sampler.hi, corpus.index_to_passage_id,
                                 num_passage_per_index_file=120)  # one file per rank: the
        # reference pairs files by substring of the end index (see refenv.digit_free_dir)
        with torch.no_grad():
            for batch in loader:
                w.add(batch["index"], model.doc(collator.encode_doc(batch["passage"])))
        w.close()
    ref.faiss_index_corpus.build_faiss_index(argparse.Namespace(index_folder=str(folder), embedding_size=DIM))
    assert sorted(os.listdir(folder)) == ["index.faiss", "index_meta.faiss"]
    indexer = ref.index.Indexer(DIM, "inner_product")
    indexer.deserialize_from(str(folder))
    assert indexer.index_id_to_db_id.tolist() == [int(d["id"]) for d in docs]
    dr = ref.retrievers.DenseRetriever(model, collator, indexer=indexer, corpus=corpus, batch_size=4)
    rng = np.random.default_rng(7)
    queries = [" ".join(rng.choice(refenv.WORDS, size=int(rng.integers(2, 12)))) for _ in range(9)]
    results = dr(queries, topk=TOPK)
    pos = {d["id"]: i for i, d in enumerate(docs)}
    D = np.array([[d["score"] for d in r] for r in results], dtype=np.float32)
    I = np.array([[pos[d["id"]] for d in r] for r in results], dtype=np.int64)
    xb = indexer.index.reconstruct_n(0, N_DOCS)
    xq = dr.calculate_query_embeddings(queries).numpy()
    assert_topk_parity(D, I, xb, xq, TOPK, what="reference DenseRetriever on the oracle")
    assert all(set(r[0].keys()) == {"id", "title", "text", "score"} for r in results)
