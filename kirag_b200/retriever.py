"""Device-resident retrieval path.

Mirror of `DenseRetriever` (/root/reference/retriever/retrievers.py:153-291) with the same
constructor arguments and the same public methods and return shapes, but:

  * query embeddings stay on the GPU between the encoder epilogue and the search — the reference
    does `.detach().cpu()` per mini-batch (retrievers.py:205), `torch.cat`, `.numpy()` (:253) and a
    host `search_knn`, i.e. two syncs and a PCIe hop per mini-batch;
  * row -> passage id mapping is one vectorised numpy take instead of the n*k `str()` loop of
    retriever/index.py:49;
  * encoding runs under `torch.no_grad()` (the reference builds an autograd graph it never uses).

`retriever`, `collator` and `corpus` are the reference's own objects (duck-typed: `.query(inputs)`,
`.encode_query(list, max_length=...)`, `.get_document(docid)`), `indexer` is a `kirag_b200.Indexer`
or the reference's `Indexer` running on `kirag_b200.as_faiss`.
"""
from __future__ import annotations

from copy import deepcopy
from typing import Dict, List, Union

import numpy as np
import torch


def _to_device(obj, device):
    if torch.is_tensor(obj):
        return obj.to(device)
    if isinstance(obj, dict):
        return {k: _to_device(v, device) for k, v in obj.items()}
    if isinstance(obj, (list, tuple)):
        return type(obj)(_to_device(v, device) for v in obj)
    return obj


class DeviceDenseRetriever(torch.nn.Module):

    def __init__(self, retriever, collator, indexer=None, corpus=None, batch_size: int = 4, **kwargs):
        super().__init__()
        self.retriever = retriever
        self.device = getattr(retriever, "device", None) or torch.device("cuda", torch.cuda.current_device())
        if hasattr(self.retriever, "eval"):
            self.retriever.eval()
        self.collator = collator
        self.indexer = indexer
        self.corpus = corpus
        self.batch_size = batch_size
        self.kwargs = kwargs

    # -- result materialisation ---------------------------------------------------------------
    def _hit(self, docid, score=None) -> dict:
        """One result entry: a private copy of the corpus document (callers mutate results) with its score, or the
        bare {"id", "score"} pair when there is no corpus (retrievers.py:265-272)."""
        if self.corpus is None:
            return {"id": docid, "score": score}
        doc = deepcopy(self.corpus.get_document(docid))
        if score is not None:
            doc["score"] = float(score)
        return doc

    def get_documents(self, docid_list: Union[List[str], Dict[str, float]]) -> List[dict]:
        """Documents for a list of ids (in order), or for an {id: score} dict (best score first, with "score")."""
        if isinstance(docid_list, dict):
            ranked = sorted(docid_list.items(), key=lambda item: item[1], reverse=True)
            return [self._hit(docid, score) for docid, score in ranked]
        if isinstance(docid_list, list):
            return [self._hit(docid) for docid in docid_list]
        raise ValueError(f"{type(docid_list)} is not a supported type for \"docid_list\"!")

    @torch.no_grad()
    def calculate_query_embeddings(self, queries: List[str], max_length: int = None, verbose: bool = False,
                                   **kwargs) -> torch.Tensor:
        """Returns a CUDA tensor [n, d] (the reference returns a CPU tensor)."""
        assert isinstance(queries, list) and len(queries) > 0  # must provide queries
        chunks = []
        for i in range(0, len(queries), self.batch_size):
            inputs = self.collator.encode_query(queries[i:i + self.batch_size], max_length=max_length, **kwargs)
            chunks.append(self.retriever.query(_to_device(inputs, self.device)).detach().float())
        return torch.cat(chunks, dim=0)

    @torch.no_grad()
    def calculate_document_embeddings(self, documents: List[str], max_length: int = None, verbose: bool = False,
                                      **kwargs) -> torch.Tensor:
        assert isinstance(documents, list) and len(documents) > 0  # must provide documents
        chunks = []
        for i in range(0, len(documents), self.batch_size):
            inputs = self.collator.encode_doc(documents[i:i + self.batch_size], max_length=max_length, **kwargs)
            chunks.append(self.retriever.doc(_to_device(inputs, self.device)).detach().float())
        return torch.cat(chunks, dim=0)

    def search_embeddings(self, queries_embeddings: torch.Tensor, topk: int):
        """[n, d] CUDA embeddings -> (passage ids as nested list of str, scores float32 ndarray [n, k])."""
        index = self.indexer.index
        D, I = index.search_device(queries_embeddings, topk)
        D, I = D.cpu().numpy(), I.cpu().numpy()  # one D2H of 12*n*k bytes
        id_map = self.indexer.index_id_to_db_id
        ids = id_map[I].astype(str).tolist()  # -1 padding maps to the last id, as in retriever/index.py:49
        return ids, D

    def parse_indexer_output(self, indexer_output) -> List[List[dict]]:
        """[(ids, scores), ...] as returned by `search_knn` -> one list of result entries per query."""
        return [[self._hit(docid, score) for docid, score in zip(ids, scores)] for ids, scores in indexer_output]

    def batch_retrieve(self, queries: List[str], topk: int, verbose: bool = False, **kwargs) -> List[dict]:
        emb = self.calculate_query_embeddings(queries=queries, verbose=verbose, **kwargs)
        ids, scores = self.search_embeddings(emb, topk)
        return self.parse_indexer_output(zip(ids, scores))

    def forward(self, queries: Union[str, List[str]], topk: int, verbose: bool = False, **kwargs):
        assert self.indexer is not None  # must provide indexer
        if isinstance(queries, str):
            return self.batch_retrieve([queries], topk=topk, verbose=verbose, **kwargs)[0]
        return self.batch_retrieve(queries, topk=topk, verbose=verbose, **kwargs)
