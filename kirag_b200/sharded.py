"""Row-sharded exact search over the GPUs of one node.

No counterpart in the reference (its FAISS search is single-process CPU,
/root/reference/retriever/index.py:47); this is the multi-GPU form
BASELINE.json's north_star asks for.  One process per GPU (torchrun).  Rank g
owns the contiguous global rows [lo_g, hi_g); every rank answers every query
against its shard with GLOBAL ids (local row + lo_g), one all-gather over
NCCL/NVLink exchanges the k per-shard candidates per query, and a merge kernel
takes the global top-k by (score desc, id asc).  Exactness: the global top-k is
a subset of the union of the per-shard top-k.

The local search and the merge are injectable so that the plumbing (shard
ranges, id offsets, gather layout) is testable with gloo on CPU, where the
tests plug the oracle in; the defaults are the CUDA library and nothing else.
"""
from __future__ import annotations

import ctypes
from typing import Callable, Optional, Tuple

import torch
import torch.distributed as dist

from . import _lib


def shard_range(n_total: int, world_size: int, rank: int) -> Tuple[int, int]:
    """Contiguous row range of `rank`: ceil-sized shards, the last ones may be short/empty."""
    per = (n_total + world_size - 1) // world_size
    lo = min(rank * per, n_total)
    hi = min(lo + per, n_total)
    return lo, hi


def merge_topk_device(D_all: torch.Tensor, I_all: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
    """D_all [G, nq, k] f32, I_all [G, nq, k] i64 (CUDA) -> (D [nq,k], I [nq,k]) via kirag_merge_topk."""
    if not D_all.is_cuda:
        raise RuntimeError("kirag_b200.sharded: merge needs CUDA tensors (no CPU fallback)")
    G, nq, k = D_all.shape
    D_all = D_all.contiguous()
    I_all = I_all.contiguous()
    D = torch.empty((nq, k), dtype=torch.float32, device=D_all.device)
    I = torch.empty((nq, k), dtype=torch.int64, device=D_all.device)
    st = torch.cuda.current_stream(D_all.device).cuda_stream
    _lib.check(
        _lib.load().kirag_merge_topk(ctypes.c_void_p(D_all.data_ptr()), ctypes.c_void_p(I_all.data_ptr()), G, nq, k,
                                     ctypes.c_void_p(D.data_ptr()), ctypes.c_void_p(I.data_ptr()), 1,
                                     D_all.device.index, ctypes.c_void_p(st)),
        "merge_topk")
    return D, I


class ShardedFlatIP:
    """SPMD sharded index: every rank calls add_shard()/search() with the same arguments."""

    def __init__(self, d: int, n_total: int, rank: Optional[int] = None, world_size: Optional[int] = None,
                 device: Optional[int] = None, group=None,
                 local_index=None, merge_fn: Optional[Callable] = None):
        self.d = d
        self.group = group
        self.rank = dist.get_rank(group) if rank is None else rank
        self.world_size = dist.get_world_size(group) if world_size is None else world_size
        self.n_total = int(n_total)
        self.lo, self.hi = shard_range(self.n_total, self.world_size, self.rank)
        if local_index is None:
            from .faiss_api import IndexFlatIP

            local_index = IndexFlatIP(d, device=device)
            local_index.reserve(max(self.hi - self.lo, 1))
        self.index = local_index
        self.merge_fn = merge_fn or merge_topk_device

    @property
    def ntotal_local(self) -> int:
        return int(self.index.ntotal)

    def add_shard(self, x_local) -> None:
        """Append rows of THIS rank's range (in order).  Tensor (CUDA) or ndarray."""
        if isinstance(x_local, torch.Tensor) and x_local.is_cuda:
            self.index.add_device(x_local.contiguous().float())
        else:
            self.index.add(x_local.numpy() if isinstance(x_local, torch.Tensor) else x_local)
        assert self.index.ntotal <= self.hi - self.lo, "more rows added than this rank's shard holds"

    def search_local(self, q: torch.Tensor, k: int) -> Tuple[torch.Tensor, torch.Tensor]:
        """Per-shard exact top-k with global ids."""
        return self.index.search_device(q, k, id_offset=self.lo)

    def search(self, q: torch.Tensor, k: int) -> Tuple[torch.Tensor, torch.Tensor]:
        """q [nq, d] replicated on every rank -> (D [nq,k], I [nq,k]) global result on every rank."""
        D_loc, I_loc = self.search_local(q, k)
        if self.world_size == 1:
            return D_loc, I_loc
        nq = q.shape[0]
        # ONE collective for scores and ids: the payload is tiny (12*nq*k bytes per rank), so the
        # exchange is latency-bound and a second all-gather would double its cost.  Scores ride in
        # the same int64 tensor as the ids (bit-cast, zero-extended).
        packed = torch.empty((2, nq, k), dtype=torch.int64, device=I_loc.device)
        packed[0] = D_loc.contiguous().view(torch.int32).to(torch.int64)
        packed[1] = I_loc
        gathered = torch.empty((self.world_size, 2, nq, k), dtype=torch.int64, device=I_loc.device)
        dist.all_gather_into_tensor(gathered.view(-1, k), packed.view(-1, k), group=self.group)
        D_all = gathered[:, 0].to(torch.int32).view(torch.float32).contiguous()
        I_all = gathered[:, 1].contiguous()
        return self.merge_fn(D_all, I_all)
