"""The reference's own callers, byte-for-byte unmodified, on the B200 library.

Everything under `ref.` below is imported from baseline/_ref — the sha256-verified snapshot of
/root/reference that `__graft_entry__.build()` takes in the build container and that travels to the GPU box
with the built .so.  `faiss` is `kirag_b200.as_faiss`.  The flow is the reference's README flow:

    compute_corpus_embeddings.cal_doc_embeddings   ->  corpus_embeddings_*.pkl / passage_id_list_*.pkl
    faiss_index_corpus.build_faiss_index           ->  index.faiss / index_meta.faiss      (faiss.write_index)
    Indexer.deserialize_from                       ->  faiss.read_index
    DenseRetriever(queries, topk)                  ->  E5Encoder -> Indexer.search_knn -> faiss IndexFlatIP.search

with a 2-layer BertConfig instance of the reference's E5Encoder / BGEEncoder (random weights, HF body) and a
word-level BertTokenizer built from an in-memory vocabulary (no network).
"""
import argparse
import os
import pickle

import numpy as np
import pytest
import torch

from tests import refenv
from tests.helpers import assert_topk_parity

pytestmark = pytest.mark.gpu

N_DOCS, DIM, TOPK = 1000, 64, 10


@pytest.fixture()
def scratch():
    with refenv.digit_free_dir() as d:
        yield d


@pytest.fixture()
def ref():
    root = refenv.snapshot_root()
    if root is None:
        pytest.skip("baseline/_ref snapshot missing or stale: run __graft_entry__.build() in the build container")
    import kirag_b200.as_faiss as af

    with refenv.reference_modules(root, af.make_module()) as ns:
        yield ns


def _world(ref, tmp_path, kind="E5Encoder", name="E5Retriever"):
    docs = refenv.synthetic_docs(N_DOCS)
    tok = refenv.tiny_tokenizer()
    refenv.save_tiny_encoder(ref, str(tmp_path / "model"), kind=kind, hidden=DIM)
    collator = (ref.collators.E5Collator if name == "E5Retriever" else ref.collators.BGECollator)(
        tokenizer=tok, query_maxlength=32, doc_maxlength=64)
    corpus = refenv.make_corpus_class(ref, docs)(title_prefix="title: ", passage_prefix="text: ")
    model = ref.retrievers.InBatchRetriever(name, str(tmp_path / "model"), local_rank=-1, temperature=0.01)
    rng = np.random.default_rng(7)
    queries = [" ".join(rng.choice(refenv.WORDS, size=int(rng.integers(2, 12)))) for _ in range(22)]
    return docs, collator, corpus, model, queries


def _results_to_arrays(results, docs):
    pos = {d["id"]: i for i, d in enumerate(docs)}
    D = np.array([[d["score"] for d in r] for r in results], dtype=np.float32)
    I = np.array([[pos[d["id"]] for d in r] for r in results], dtype=np.int64)
    return D, I


def _build_with_reference_scripts(ref, scratch, model, corpus, collator):
    args = argparse.Namespace(local_rank=-1, save_dir=scratch, name="run", index_folder="idx",
                              num_passage_per_index_file=300, embedding_size=DIM)
    loader = ref.utils.get_dataloader(-1, corpus, 8, shuffle=False, drop_last=False)
    ref.compute_corpus_embeddings.cal_doc_embeddings(args, model, loader, collator)
    folder = os.path.join(scratch, "run", "idx")
    written = sorted(os.listdir(folder))
    assert written == sorted([f"corpus_embeddings_{s}_{e}.pkl" for s, e in ((0, 303), (304, 607), (608, 911), (912, 999))] +
                             [f"passage_id_list_{s}_{e}.pkl" for s, e in ((0, 303), (304, 607), (608, 911), (912, 999))])
    ref_files = {f: pickle.load(open(os.path.join(folder, f), "rb")) for f in written}
    args.index_folder = folder
    ref.faiss_index_corpus.build_faiss_index(args)
    assert sorted(os.listdir(folder)) == ["index.faiss", "index_meta.faiss"]
    indexer = ref.index.Indexer(DIM, "inner_product")
    indexer.deserialize_from(folder)
    return indexer, ref_files, folder


def test_reference_readme_flow_unmodified_on_the_b200_library(ref, tmp_path, scratch):
    from kirag_b200 import faiss_api

    docs, collator, corpus, model, queries = _world(ref, tmp_path)
    indexer, ref_files, folder = _build_with_reference_scripts(ref, scratch, model, corpus, collator)
    assert isinstance(indexer.index, faiss_api.IndexFlatIP) and indexer.index.ntotal == N_DOCS
    assert indexer.index_id_to_db_id.tolist() == [int(d["id"]) for d in docs]
    dr = ref.retrievers.DenseRetriever(model, collator, indexer=indexer, corpus=corpus, batch_size=4)
    results = dr(queries, topk=TOPK)
    assert len(results) == len(queries) and all(len(r) == TOPK for r in results)
    assert all(set(r[0].keys()) == {"id", "title", "text", "score"} for r in results)
    D, I = _results_to_arrays(results, docs)
    xb = indexer.index.reconstruct_n(0, N_DOCS)
    xq = dr.calculate_query_embeddings(queries).numpy()
    np.testing.assert_allclose(np.linalg.norm(xb, axis=1), 1.0, atol=1e-5)
    n_swaps = assert_topk_parity(D, I, xb, xq, TOPK, what="reference DenseRetriever on as_faiss")
    assert n_swaps <= 2
    # the same answer as the reference's arithmetic on the stored embeddings (torch fp32 on the host)
    sims = torch.from_numpy(xq) @ torch.from_numpy(xb).T
    ts, ti = torch.topk(sims, TOPK, dim=1)
    np.testing.assert_allclose(D, ts.numpy(), rtol=1e-5, atol=1e-6)
    assert (I != ti.numpy()).sum() <= 2
    # a single str query returns one list (retrievers.py:288-289); no corpus -> {"id","score"} dicts (:271-272)
    one = dr(queries[3], topk=5)
    assert [d["id"] for d in one] == [d["id"] for d in results[3][:5]]
    bare = ref.retrievers.DenseRetriever(model, collator, indexer=indexer, corpus=None, batch_size=4)
    assert set(bare([queries[0]], topk=3)[0][0].keys()) == {"id", "score"}
    # stored embeddings == what the reference pickled (index.add copies fp32 verbatim)
    emb = torch.cat([ref_files[f"corpus_embeddings_{s}_{e}.pkl"] for s, e in ((0, 303), (304, 607), (608, 911), (912, 999))])
    assert np.array_equal(emb.numpy(), xb)


@pytest.mark.parametrize("kind,name", [("E5Encoder", "E5Retriever"), ("BGEEncoder", "BGERetriever")])
def test_patched_reference_encoders_match_the_unpatched_forward(ref, tmp_path, kind, name):
    """retriever/encoders.py:67-77,106-118 with the tail replaced by the fused epilogue kernel
    (pooling.patch_reference_encoders): forward in fp32 and under bf16 autocast, and the gradients of
    InBatchRetriever.forward (retrievers.py:133-150) through the kernel's analytic backward."""
    from kirag_b200 import pooling

    docs, collator, corpus, model, queries = _world(ref, tmp_path, kind=kind, name=name)
    model = model.cuda()
    q_in = ref.utils.to_device(collator.encode_query(queries[:8]), model.device)
    d_in = ref.utils.to_device(collator.encode_doc([corpus[i]["passage"] for i in range(16)]), model.device)
    labels = torch.arange(8, device=model.device) * 2

    def run():
        model.eval()
        with torch.no_grad():
            e32 = model.query(q_in).float().cpu()
            with torch.autocast("cuda", dtype=torch.bfloat16):
                e16 = model.query(q_in).float().cpu()
        model.train()
        model.zero_grad()
        loss, scores, _, _ = model(q_in, d_in, labels=labels)
        loss.backward()
        grads = {n: p.grad.detach().float().cpu().clone() for n, p in model.named_parameters() if p.grad is not None}
        return e32, e16, float(loss), grads

    base = run()
    undo = pooling.patch_reference_encoders(ref.encoders)
    try:
        fused = run()
    finally:
        undo()
    again = run()  # un-patching restores the reference's own forward
    assert torch.equal(again[0], base[0])
    assert float((fused[0] - base[0]).abs().max()) < 1e-5, "fp32 forward"
    assert float((fused[1] - base[1]).abs().max()) < 1e-3, "bf16 autocast forward"  # north_star tolerance
    assert abs(fused[2] - base[2]) < 1e-3 * max(1.0, abs(base[2]))
    assert fused[3].keys() == base[3].keys() and len(base[3]) > 10
    # gradients: fp32 noise of a loss with temperature 0.01 is amplified on parameters whose gradient is a small
    # sum with cancellation, so the bound is relative to the whole gradient vector, plus a loose per-tensor one
    num = sum(float(((fused[3][n] - base[3][n]) ** 2).sum()) for n in base[3])
    den = sum(float((base[3][n] ** 2).sum()) for n in base[3])
    assert den > 0 and (num / den) ** 0.5 < 1e-3, (num, den)
    for n in base[3]:
        scale = float(base[3][n].abs().max()) + 1e-6
        assert float((fused[3][n] - base[3][n]).abs().max()) <= 2e-2 * scale + 1e-6, n


def test_device_resident_pipeline_equals_the_reference_host_path(ref, tmp_path, scratch):
    """§8f n2: encoder -> pool_normalize_kernel -> search_device with no host hop, against the reference's
    DenseRetriever (.cpu() per mini-batch, numpy, host search_knn) on the same index."""
    from kirag_b200 import pooling
    from kirag_b200.retriever import DeviceDenseRetriever

    docs, collator, corpus, model, queries = _world(ref, tmp_path)
    indexer, _, _ = _build_with_reference_scripts(ref, scratch, model, corpus, collator)
    host = ref.retrievers.DenseRetriever(model, collator, indexer=indexer, corpus=corpus, batch_size=4)
    want = host(queries, topk=TOPK)
    undo = pooling.patch_reference_encoders(ref.encoders)
    try:
        dev = DeviceDenseRetriever(model, collator, indexer=indexer, corpus=corpus, batch_size=4)
        emb = dev.calculate_query_embeddings(queries)
        assert emb.is_cuda and emb.shape == (len(queries), DIM)
        got = dev(queries, topk=TOPK)
    finally:
        undo()
    Dw, Iw = _results_to_arrays(want, docs)
    Dg, Ig = _results_to_arrays(got, docs)
    np.testing.assert_allclose(Dg, Dw, rtol=1e-5, atol=2e-6)
    assert (Ig != Iw).sum() <= 2  # fp32 near-ties only
    assert [set(d.keys()) for d in got[0]] == [set(d.keys()) for d in want[0]]


def test_embedding_writer_with_a_real_encoder_equals_the_reference_producer(ref, tmp_path, scratch):
    """§8 a12 / n4: the per-rank contiguous writer driven by the reference's encoder with the fused epilogue
    produces the embeddings `cal_doc_embeddings` (compute_corpus_embeddings.py:50-134) pickles, and
    kirag_b200.build_index turns them into the same index."""
    from kirag_b200 import build_index, pooling
    from kirag_b200.embed_writer import ContiguousShardSampler, EmbeddingShardWriter

    docs, collator, corpus, model, queries = _world(ref, tmp_path)
    indexer, ref_files, _ = _build_with_reference_scripts(ref, scratch, model, corpus, collator)
    xb_ref = indexer.index.reconstruct_n(0, N_DOCS)
    out = tmp_path / "ours"
    model = model.cuda().eval()
    undo = pooling.patch_reference_encoders(ref.encoders)
    try:
        for rank in range(2):  # two ranks' ranges, written one after the other by this process
            sampler = ContiguousShardSampler(len(corpus), rank, 2)
            loader = torch.utils.data.DataLoader(corpus, batch_size=8, sampler=sampler)
            w = EmbeddingShardWriter(str(out), DIM, sampler.lo, sampler.hi, corpus.index_to_passage_id,
                                     num_passage_per_index_file=300)
            with torch.no_grad():
                for batch in loader:
                    args = ref.utils.to_device(collator.encode_doc(batch["passage"]), model.device)
                    w.add(batch["index"], model.doc(args))
            w.close()
    finally:
        undo()
    pairs = build_index.pair_embedding_files(str(out))
    assert [os.path.basename(a) for a, _ in pairs] == [f"corpus_embeddings_{s}_{e}.pkl" for s, e in
                                                       ((0, 299), (300, 499), (500, 599), (600, 899), (900, 999))]
    ours = build_index.build_faiss_index(index_folder=str(out), embedding_size=DIM)
    xb = ours.index.reconstruct_n(0, N_DOCS)
    assert ours.index_id_to_db_id.tolist() == indexer.index_id_to_db_id.tolist()
    assert float(np.abs(xb - xb_ref).max()) < 1e-5
    xq = np.ascontiguousarray(xb_ref[::50])
    a = indexer.search_knn(xq, TOPK, verbose=False)
    b = ours.search_knn(xq, TOPK, verbose=False)
    assert sum(x[0] != y[0] for x, y in zip(a, b)) <= 1
