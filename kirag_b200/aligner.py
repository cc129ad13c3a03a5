"""Reasoning-Chain-Aligner triple scoring with cached triple embeddings.

`KiRAG.filter_candidate_triples` (/root/reference/knowledge_graph/models.py:1514-1542) re-embeds
ALL accumulated candidate triples on every reasoning turn (batch 4, `.cpu()` per batch,
retrievers.py:214-232) and multiplies on the CPU.  `TripleScorer` keeps the embedding of every triple
text it has seen on the GPU, embeds only the new ones, and runs the same contraction + top-k through
the search kernel (`kirag_topk_ip`).  Return shape is the reference's: (indices, scores) nested lists.
"""
from __future__ import annotations

from typing import Callable, Dict, List, Tuple

import torch

from .scoring import topk_inner_product


class TripleScorer:

    def __init__(self, embed_queries: Callable[[List[str]], torch.Tensor],
                 embed_documents: Callable[[List[str]], torch.Tensor], max_cached: int = 1_000_000):
        """embed_*: callables returning [n, d] embeddings (CUDA or CPU tensors), e.g. the aligner's
        `calculate_query_embeddings(queries=..., max_length=256)` / `calculate_document_embeddings(...)`."""
        self.embed_queries = embed_queries
        self.embed_documents = embed_documents
        self.max_cached = max_cached
        self._row_of: Dict[str, int] = {}
        self._bank = None  # [n_cached, d] float32 CUDA
        self.n_embedded = 0  # triples actually sent through the encoder (for tests / accounting)

    def _ensure(self, texts: List[str]) -> torch.Tensor:
        """Embeddings [len(texts), d] (CUDA) of `texts`, embedding only what the bank does not hold yet."""
        wanted = list(dict.fromkeys(texts))
        missing = [t for t in wanted if t not in self._row_of]
        held = 0 if self._bank is None else self._bank.shape[0]
        if missing and held + len(missing) > self.max_cached:
            # the bank is full: start over with exactly what this call needs (everything is re-embedded once)
            self._row_of.clear()
            self._bank, held = None, 0
            missing = wanted
        if missing:
            emb = self.embed_documents(missing)
            if not emb.is_cuda:
                emb = emb.cuda()
            emb = emb.detach().float()
            self.n_embedded += len(missing)
            self._bank = emb if self._bank is None else torch.cat([self._bank, emb], dim=0)
            for i, t in enumerate(missing):
                self._row_of[t] = held + i
        rows = [self._row_of[t] for t in texts]
        if rows == list(range(self._bank.shape[0])):
            return self._bank  # the candidates ARE the bank, in order: no gather
        if rows and rows == list(range(rows[0], rows[0] + len(rows))):
            return self._bank[rows[0]:rows[0] + len(rows)]  # a contiguous run: a view
        return self._bank.index_select(0, torch.tensor(rows, dtype=torch.int64, device=self._bank.device))

    def filter_candidate_triples(self, query_texts: List[str], triple_texts: List[str],
                                 num_candidate_triples: int) -> Tuple[List[List[int]], List[List[float]]]:
        """query_texts: the "{question}\\nknowledge triples: ..." strings of models.py:1526;
        triple_texts: `get_triple_text(triple)` of every candidate.  Returns (indices, scores)."""
        q = self.embed_queries(query_texts)
        if not q.is_cuda:
            q = q.cuda()
        t = self._ensure(triple_texts)
        D, I = topk_inner_product(q.detach().float(), t, num_candidate_triples)
        return I.tolist(), D.tolist()
