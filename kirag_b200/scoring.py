"""Aligner triple scoring on the search kernel.

Drop-in for the arithmetic of KiRAG.filter_candidate_triples
(/root/reference/knowledge_graph/models.py:1532-1542):

    sims = torch.matmul(queries_embeddings, triples_embeddings.T)
    scores, indices = torch.topk(sims, k=min(num_candidate_triples, num_triples), dim=1)
    return indices.tolist(), scores.tolist()

The reference does this on CPU tensors; here the same contraction + top-k is
one call into the B200 library (the flat-IP search over a transient matrix).
Ties are ordered by lower index (torch.topk leaves tie order unspecified).
"""
from __future__ import annotations

import ctypes
from typing import List, Tuple

import numpy as np

from . import _lib


def topk_inner_product(queries, triples, k: int):
    """queries [C, d], triples [T, d] (numpy or torch, host or CUDA) -> (scores [C,k'], indices [C,k']),
    k' = min(k, T), as numpy arrays (host inputs) or CUDA tensors (CUDA inputs)."""
    lib = _lib.load()
    is_torch = hasattr(queries, "is_cuda")
    if is_torch and queries.is_cuda:
        import torch

        q = queries.contiguous().float()
        t = triples.to(q.device).contiguous().float()
        # the kernels use 16-byte loads: a view whose storage offset breaks that alignment is copied
        if q.data_ptr() % 16:
            q = q.clone()
        if t.data_ptr() % 16:
            t = t.clone()
        C, d = q.shape
        T = t.shape[0]
        kk = min(int(k), T)
        assert kk > 0, "need at least one candidate triple and k > 0"
        D = torch.empty((C, kk), dtype=torch.float32, device=q.device)
        I = torch.empty((C, kk), dtype=torch.int64, device=q.device)
        st = torch.cuda.current_stream(q.device).cuda_stream
        _lib.check(lib.kirag_topk_ip(ctypes.c_void_p(q.data_ptr()), C, ctypes.c_void_p(t.data_ptr()), T, d, kk,
                                     ctypes.c_void_p(D.data_ptr()), ctypes.c_void_p(I.data_ptr()), 1,
                                     q.device.index, ctypes.c_void_p(st)), "topk_ip")
        return D, I
    q = np.ascontiguousarray(queries.numpy() if is_torch else queries, dtype=np.float32)
    t = triples.numpy() if hasattr(triples, "numpy") else triples
    t = np.ascontiguousarray(t, dtype=np.float32)
    C, d = q.shape
    T = t.shape[0]
    kk = min(int(k), T)
    assert kk > 0, "need at least one candidate triple and k > 0"
    D = np.empty((C, kk), dtype=np.float32)
    I = np.empty((C, kk), dtype=np.int64)
    _lib.check(lib.kirag_topk_ip(ctypes.c_void_p(q.ctypes.data), C, ctypes.c_void_p(t.ctypes.data), T, d, kk,
                                 ctypes.c_void_p(D.ctypes.data), ctypes.c_void_p(I.ctypes.data), 0,
                                 _lib.default_device(), None), "topk_ip")
    return D, I


def filter_candidate_triples_scores(queries_embeddings, triples_embeddings,
                                    num_candidate_triples: int) -> Tuple[List[List[int]], List[List[float]]]:
    """Returns (indices, scores) as nested lists, exactly the return shape of models.py:1539-1542."""
    D, I = topk_inner_product(queries_embeddings, triples_embeddings, num_candidate_triples)
    return I.tolist(), D.tolist()
