"""Helpers that run the reference's OWN modules (unmodified) inside a test.

The reference has no plugin registry: its seam is `import faiss` in retriever/index.py:6.  `reference_modules`
puts a faiss-shaped module into sys.modules (the CUDA library's stand-in on the GPU box, the oracle's on the
CPU), makes the reference importable from a root directory and undoes both afterwards.  The root is
`baseline/_ref` (the snapshot `__graft_entry__.build()` takes, see tools/snapshot_reference.py) — the only copy
that exists on the GPU box — or /root/reference in the build container.
"""
from __future__ import annotations

import contextlib
import importlib
import os
import sys
import types

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SNAPSHOT = os.path.join(ROOT, "baseline", "_ref")
REFERENCE = "/root/reference"
_REF_TOPLEVEL = ("retriever", "dataset", "utils", "compute_corpus_embeddings", "faiss_index_corpus")


def snapshot_root():
    """baseline/_ref if the snapshot is there and intact, else None."""
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    try:
        import snapshot_reference
    finally:
        sys.path.pop(0)
    return SNAPSHOT if snapshot_reference.verify(SNAPSHOT) else None


def any_reference_root():
    """The snapshot, else /root/reference (build container), else None."""
    root = snapshot_root()
    if root:
        return root
    return REFERENCE if os.path.isdir(os.path.join(REFERENCE, "retriever")) else None


def _is_ref_module(name: str) -> bool:
    return name.split(".")[0] in _REF_TOPLEVEL


@contextlib.contextmanager
def reference_modules(root: str, faiss_module: types.ModuleType):
    saved = {k: v for k, v in sys.modules.items() if _is_ref_module(k) or k == "faiss"}
    for k in list(saved):
        sys.modules.pop(k, None)
    sys.modules["faiss"] = faiss_module
    sys.path.insert(0, root)
    try:
        ns = types.SimpleNamespace(
            index=importlib.import_module("retriever.index"),
            encoders=importlib.import_module("retriever.encoders"),
            retrievers=importlib.import_module("retriever.retrievers"),
            collators=importlib.import_module("dataset.collators"),
            corpus=importlib.import_module("dataset.corpus"),
            utils=importlib.import_module("utils.utils"),
            faiss_index_corpus=importlib.import_module("faiss_index_corpus"),
            compute_corpus_embeddings=importlib.import_module("compute_corpus_embeddings"),
        )
        assert os.path.realpath(ns.index.__file__).startswith(os.path.realpath(root)), ns.index.__file__
        yield ns
    finally:
        sys.path.remove(root)
        for k in [k for k in sys.modules if _is_ref_module(k) or k == "faiss"]:
            sys.modules.pop(k, None)
        sys.modules.update(saved)


# ------------------------------------------------------------------ tiny synthetic world ---
WORDS = [f"w{i}" for i in range(400)]
VOCAB = ["[PAD]", "[UNK]", "[CLS]", "[SEP]", "[MASK]", "query", "passage", "title", "text", ":", ","] + WORDS


def tiny_tokenizer():
    from transformers import BertTokenizer

    return BertTokenizer(vocab={w: i for i, w in enumerate(VOCAB)})


def save_tiny_encoder(ref, path: str, kind: str = "E5Encoder", hidden: int = 64, seed: int = 0) -> None:
    """A randomly initialised 2-layer BertConfig instance of the reference's own encoder class, saved so that
    the reference's `load_retriever` (`from_pretrained`, retrievers.py:25-29) can load it."""
    import torch
    from transformers import BertConfig

    cfg = BertConfig(hidden_size=hidden, num_hidden_layers=2, num_attention_heads=4, intermediate_size=2 * hidden,
                     vocab_size=len(VOCAB), max_position_embeddings=96, hidden_dropout_prob=0.0,
                     attention_probs_dropout_prob=0.0)
    torch.manual_seed(seed)
    getattr(ref.encoders, kind)(cfg).save_pretrained(path)


def synthetic_docs(n: int, seed: int = 0):
    import numpy as np

    rng = np.random.default_rng(seed)
    docs = []
    for i in range(n):
        title = " ".join(rng.choice(WORDS, size=int(rng.integers(1, 4))))
        text = " ".join(rng.choice(WORDS, size=int(rng.integers(3, 40))))
        docs.append({"id": str(5000 + 3 * i), "title": title, "text": text})
    return docs


def make_corpus_class(ref, docs):
    """A corpus in the reference's own style (dataset/corpus.py: one subclass per dataset)."""

    class SyntheticCorpus(ref.corpus.Corpus):
        def load_corpus_data(self):
            return docs

        def doc_to_str(self, doc):
            return self.passage_format.format(title_prefix=self.title_prefix, title=doc["title"],
                                              passage_prefix=self.passage_prefix, passage=doc["text"]).strip()

        def __getitem__(self, index):
            ex = self.data[index]
            return {"index": index, "passage_id": ex["id"], "passage": self.doc_to_str(ex)}

    return SyntheticCorpus


@contextlib.contextmanager
def digit_free_dir():
    """A scratch directory whose PATH contains no digits.  The reference pairs embedding and passage-id files by
    SUBSTRING of the whole path (faiss_index_corpus.py:37-41: `if embedding_end_passage_id_str in passage_id_file`)
    and parses the end index with `split(".")[0]` (:24), so a pytest tmp_path such as .../pytest-30/... can pair
    every file with the wrong partner.  Running the reference unmodified means giving it a path it can handle."""
    import shutil
    import tempfile

    tag = "".join(chr(ord("a") + int(c)) for c in str(os.getpid()))
    base = os.path.join(tempfile.gettempdir(), "kiragrefflow_" + tag)
    assert not any(c.isdigit() for c in base) and "." not in base, base
    shutil.rmtree(base, ignore_errors=True)
    os.makedirs(base)
    try:
        yield base
    finally:
        shutil.rmtree(base, ignore_errors=True)
