"""Corpus-embedding writer.

The reference's `cal_doc_embeddings` (/root/reference/compute_corpus_embeddings.py:77-125) gathers
every batch to rank 0 (`dist.gather` x2 + two barriers per step, utils/utils.py:145-155), grows a CPU
tensor with `torch.cat` per batch (O(n^2) copies, :94-97) and re-orders through a Python dict before
pickling.  Here every rank encodes a CONTIGUOUS range of the corpus and writes its own files, with no
collective at all on the data path:

  ContiguousShardSampler  rank r iterates corpus rows [lo_r, hi_r) in order (instead of the strided
                          DistributedSampler of utils/utils.py:120-122)
  EmbeddingShardWriter    rows land in a pre-allocated buffer (on whatever device the embeddings are
                          on); a file pair is written whenever `num_passage_per_index_file` rows are
                          complete, in the reference's own format and naming
                          (corpus_embeddings_{s}_{e}.pkl = CPU FloatTensor, passage_id_list_{s}_{e}.pkl
                          = list of passage ids, compute_corpus_embeddings.py:114-115),
                          so `faiss_index_corpus.build_faiss_index` / `kirag_b200.build_index` consume
                          them unchanged.
"""
from __future__ import annotations

import os
import pickle
from typing import Iterator, List, Optional, Sequence

import torch

from .sharded import shard_range


class ContiguousShardSampler(torch.utils.data.Sampler):
    """Rank `rank` of `world_size` visits dataset rows [lo, hi) in ascending order."""

    def __init__(self, n: int, rank: int = 0, world_size: int = 1):
        self.n, self.rank, self.world_size = int(n), int(rank), int(world_size)
        self.lo, self.hi = shard_range(self.n, self.world_size, self.rank)

    def __iter__(self) -> Iterator[int]:
        return iter(range(self.lo, self.hi))

    def __len__(self) -> int:
        return self.hi - self.lo


class EmbeddingShardWriter:
    def __init__(self, save_dir: str, dim: int, start: int, end: int, index_to_passage_id: Sequence,
                 num_passage_per_index_file: int = 1_000_000, device=None):
        """Writes the embeddings of corpus rows [start, end) (this rank's range).  File boundaries are
        global multiples of `num_passage_per_index_file`, so ranks never produce overlapping names."""
        self.save_dir, self.dim = save_dir, int(dim)
        self.start, self.end = int(start), int(end)
        self.per_file = int(num_passage_per_index_file)
        self.ids = index_to_passage_id
        self.device = device
        os.makedirs(save_dir, exist_ok=True)
        self._buf: Optional[torch.Tensor] = None
        self._filled: Optional[torch.Tensor] = None
        self._chunk_lo = self.start  # first row of the chunk being filled
        self.files: List[str] = []

    def _chunk_hi(self, lo: int) -> int:
        return min((lo // self.per_file + 1) * self.per_file, self.end)

    def _alloc(self, like: torch.Tensor) -> None:
        rows = self._chunk_hi(self._chunk_lo) - self._chunk_lo
        self._buf = torch.empty((rows, self.dim), dtype=torch.float32, device=like.device)
        self._filled = torch.zeros((rows,), dtype=torch.bool, device=like.device)

    def add(self, corpus_indices, embeddings: torch.Tensor) -> None:
        """corpus_indices: int tensor/list [b] of global corpus rows; embeddings: [b, dim]."""
        idx = torch.as_tensor(corpus_indices, dtype=torch.int64, device=embeddings.device)
        emb = embeddings.detach().float()
        assert emb.dim() == 2 and emb.shape[1] == self.dim and idx.shape[0] == emb.shape[0]
        assert int(idx.min()) >= self.start and int(idx.max()) < self.end, "row outside this writer's range"
        while idx.numel():
            if self._buf is None:
                self._alloc(emb)
            hi = self._chunk_hi(self._chunk_lo)
            here = (idx >= self._chunk_lo) & (idx < hi)
            if bool(here.any()):
                local = idx[here] - self._chunk_lo
                self._buf.index_copy_(0, local, emb[here])
                self._filled[local] = True
            rest = idx >= hi
            assert not bool((idx < self._chunk_lo).any()), "rows must arrive chunk by chunk (use ContiguousShardSampler)"
            if bool(self._filled.all()):
                self._flush()
            elif bool(rest.any()):
                raise AssertionError("rows of the next file arrived before the current one is complete "
                                     "(use ContiguousShardSampler)")
            idx, emb = idx[rest], emb[rest]

    def _flush(self) -> None:
        lo, hi = self._chunk_lo, self._chunk_lo + self._buf.shape[0]
        emb_file = os.path.join(self.save_dir, f"corpus_embeddings_{lo}_{hi - 1}.pkl")
        ids_file = os.path.join(self.save_dir, f"passage_id_list_{lo}_{hi - 1}.pkl")
        with open(emb_file, "wb") as f:
            pickle.dump(self._buf.cpu(), f)
        with open(ids_file, "wb") as f:
            pickle.dump([self.ids[i] for i in range(lo, hi)], f)
        self.files += [emb_file, ids_file]
        self._chunk_lo = hi
        self._buf = self._filled = None

    def close(self) -> List[str]:
        if self._buf is not None:
            missing = int((~self._filled).sum())
            assert missing == 0, f"{missing} rows of chunk starting at {self._chunk_lo} were never added"
        assert self._chunk_lo == self.end or self.start == self.end, \
            f"rows [{self._chunk_lo}, {self.end}) were never added"
        return self.files
