"""`Indexer` / `ShardedIndexer`: the reference's indexer interface on the B200 library.

The reference's own `retriever/index.py` runs unmodified on `kirag_b200.as_faiss` (tests/test_reference_*_gpu.py);
this module is for callers that want an indexer object WITHOUT the reference on their path.  `Indexer` offers the
public surface of /root/reference/retriever/index.py:17-83 — `index_data`, `search_knn`, `serialize`,
`deserialize_from`, the attributes `index` and `index_id_to_db_id`, the file names `index.faiss` /
`index_meta.faiss` and the return shape `[(ids: list[str], scores: float32[k]), ...]` — and is written around what
differs from the reference's glue:

  * the row -> passage-id table is an amortised-growth int64 array (the reference re-concatenates the whole table
    on every `index_data`, index.py:81-83: O(total) per call);
  * row numbers are turned into passage-id strings with ONE vectorised take + `astype(str)` per faiss call (the
    reference runs a Python `str()` per result, index.py:49: 1.6M calls at 16384 queries x 100);
  * only the exact inner-product index exists here (the only one any KiRAG caller constructs: retrieve.py:112,
    faiss_index_corpus.py:29); asking for "l2" or product quantisation fails at construction.
"""
from __future__ import annotations

import logging
import os
import pickle
from typing import Iterable, List, Sequence, Tuple

import numpy as np

from . import faiss_api

logger = logging.getLogger(__name__)

INDEX_FILE = "index.faiss"      # index.py:57,68
META_FILE = "index_meta.faiss"  # index.py:58,69
# metric name -> index class; tests substitute a CPU test double here
INDEX_TYPES = {"inner_product": faiss_api.IndexFlatIP}


class _PassageIds:
    """Append-only int64 table: index row -> passage id, with amortised growth."""

    def __init__(self, initial: Sequence[int] = ()):
        self._buf = np.array(initial, dtype=np.int64).reshape(-1)
        self._n = self._buf.shape[0]

    def extend(self, ids: Iterable) -> None:
        new = np.asarray(list(ids) if not isinstance(ids, np.ndarray) else ids)
        new = new.astype(np.int64).reshape(-1)  # "123" -> 123, like np.array(db_ids, dtype=np.int64) in the reference
        need = self._n + new.shape[0]
        if need > self._buf.shape[0]:
            grown = np.empty(max(need, 2 * self._buf.shape[0], 1024), dtype=np.int64)
            grown[:self._n] = self._buf[:self._n]
            self._buf = grown
        self._buf[self._n:need] = new
        self._n = need

    def view(self) -> np.ndarray:
        return self._buf[:self._n]

    def __len__(self) -> int:
        return self._n


class Indexer:

    def __init__(self, vector_sz, metric="inner_product", n_subquantizers=0, n_bits=8, device=None):
        if n_subquantizers > 0:
            raise NotImplementedError("product quantisation is not part of this library (no KiRAG caller uses it)")
        if metric not in INDEX_TYPES:
            raise NotImplementedError(f"metric {metric!r}: only 'inner_product' is implemented "
                                      "(the only metric KiRAG constructs)")
        make = INDEX_TYPES[metric]
        self.index = make(int(vector_sz)) if device is None else make(int(vector_sz), device=device)
        self._ids = _PassageIds()

    # the reference exposes the table as a plain attribute; keep it readable and assignable
    @property
    def index_id_to_db_id(self) -> np.ndarray:
        return self._ids.view()

    @index_id_to_db_id.setter
    def index_id_to_db_id(self, table) -> None:
        self._ids = _PassageIds(np.asarray(table, dtype=np.int64))

    def index_data(self, ids, embeddings) -> None:
        """Append rows `embeddings` [n, d] (any float dtype) whose passage ids are `ids` (int-parsable)."""
        rows = np.ascontiguousarray(embeddings, dtype=np.float32)
        if len(ids) != rows.shape[0]:
            raise AssertionError(f"{len(ids)} ids for {rows.shape[0]} embeddings")
        self._ids.extend(ids)
        self.index.add(rows)
        logger.info("Total data indexed %d", len(self._ids))

    def _rows_to_ids(self, rows: np.ndarray) -> List[List[str]]:
        table = self._ids.view()
        if table.shape[0] == 0:
            return rows.astype(str).tolist()
        # numpy's negative indexing maps FAISS's -1 padding to the LAST id, which is what the reference's
        # `self.index_id_to_db_id[i]` does with it (index.py:49)
        return table[rows].astype(str).tolist()

    def search_knn(self, query_vectors, top_docs: int, index_batch_size: int = 1024,
                   verbose: bool = True) -> List[Tuple[List[str], np.ndarray]]:
        queries = np.ascontiguousarray(query_vectors, dtype=np.float32)
        out: List[Tuple[List[str], np.ndarray]] = []
        for lo in range(0, queries.shape[0], int(index_batch_size)):
            scores, rows = self.index.search(queries[lo:lo + int(index_batch_size)], top_docs)
            out.extend(zip(self._rows_to_ids(rows), scores))
        return out

    def serialize(self, dir_path) -> None:
        index_file, meta_file = os.path.join(dir_path, INDEX_FILE), os.path.join(dir_path, META_FILE)
        logger.info("Serializing index to %s, meta data to %s", index_file, meta_file)
        faiss_api.write_index(self.index, index_file)
        with open(meta_file, "wb") as f:
            pickle.dump(np.array(self._ids.view()), f)  # a plain int64 ndarray, as the reference pickles

    def deserialize_from(self, dir_path) -> None:
        index_file, meta_file = os.path.join(dir_path, INDEX_FILE), os.path.join(dir_path, META_FILE)
        logger.info("Loading index from %s, meta data from %s", index_file, meta_file)
        self.index = faiss_api.read_index(index_file, faiss_api.IO_FLAG_MMAP)
        with open(meta_file, "rb") as f:
            self.index_id_to_db_id = pickle.load(f)
        if len(self._ids) != self.index.ntotal:
            raise AssertionError(f"{meta_file} holds {len(self._ids)} passage ids, {index_file} {self.index.ntotal} rows")


class ShardedIndexer(object):
    """`Indexer` over the GPUs of one node (one process per GPU, SPMD): same five methods and return types
    as /root/reference/retriever/index.py:17-83, every rank makes the same calls with the same arguments and
    gets the same results.

    The reference has no multi-GPU index (its FAISS search is single-process CPU).  `index_data(ids,
    embeddings)` is called with the SAME chunk on every rank (as `build_faiss_index` iterates over the saved
    shards, faiss_index_corpus.py:42-46); chunk number c is kept by rank c % world_size, so no row count has
    to be known in advance.  Each rank answers a query batch against its rows, maps local rows to passage ids
    ON THE DEVICE, and the per-rank (score, passage id) lists are exchanged and merged by the fused peer-memory
    kernel (kirag_b200.sharded).  Ties are broken by the lower passage id (the reference's ids are the row
    numbers of the corpus file, preprocessing/…:67, so this is FAISS's lower-row rule).
    """

    def __init__(self, vector_sz, metric="inner_product", n_subquantizers=0, n_bits=8, device=None, rank=None,
                 world_size=None, group=None, local_index=None, merge_fn=None, exchange=None, max_nq=1024, max_k=128):
        import torch.distributed as dist

        from .sharded import ShardedFlatIP

        assert metric == "inner_product" and n_subquantizers == 0, "only the flat inner-product index is sharded"
        self.rank = dist.get_rank(group) if rank is None else rank
        self.world_size = dist.get_world_size(group) if world_size is None else world_size
        # n_total is unknown up front: the row-range bookkeeping of ShardedFlatIP is not used, only its local
        # index and its exchange
        self._sh = ShardedFlatIP(vector_sz, 0, rank=self.rank, world_size=self.world_size, device=device, group=group,
                                 local_index=local_index, merge_fn=merge_fn, exchange=exchange, max_nq=max_nq, max_k=max_k)
        self.index = self._sh.index
        self.index_id_to_db_id = np.empty((0), dtype=np.int64)  # passage ids of THIS rank's rows
        self._ids_dev = None
        self._chunks_seen = 0
        self.ntotal_global = 0
        # the sort-free merge of the peer kernel needs every per-rank list ordered by (score desc, id asc): true
        # as long as passage ids increase with the row number (decided on the ids EVERY rank sees, so that all
        # ranks take the same path)
        self._ids_increasing = True
        self._last_id = None

    def index_data(self, ids, embeddings):
        new_ids = np.array(ids, dtype=np.int64)
        if len(new_ids):
            if (self._last_id is not None and new_ids[0] <= self._last_id) or np.any(new_ids[1:] <= new_ids[:-1]):
                self._ids_increasing = False
            self._last_id = int(new_ids[-1])
        mine = (self._chunks_seen % self.world_size) == self.rank
        self._chunks_seen += 1
        self.ntotal_global += len(ids)
        if not mine:
            return
        self.index_id_to_db_id = np.concatenate((self.index_id_to_db_id, new_ids), axis=0)
        self._ids_dev = None
        self.index.add(np.ascontiguousarray(embeddings, dtype=np.float32) if isinstance(embeddings, np.ndarray)
                       else embeddings.astype('float32'))

    def _local_ids_on(self, device):
        import torch

        if self._ids_dev is None or self._ids_dev.device != device:
            self._ids_dev = torch.from_numpy(self.index_id_to_db_id).to(device)
        return self._ids_dev

    def search_knn(self, query_vectors: np.array, top_docs: int, index_batch_size=1024,
                   verbose: bool = True) -> List[Tuple[List[object], List[float]]]:
        import torch

        query_vectors = np.ascontiguousarray(query_vectors, dtype=np.float32)
        result = []
        nbatch = (len(query_vectors) - 1) // index_batch_size + 1
        for b in range(nbatch):
            q_np = query_vectors[b * index_batch_size:min((b + 1) * index_batch_size, len(query_vectors))]
            q = torch.from_numpy(q_np)
            dev = getattr(self.index, "device", None)
            if dev is not None:
                q = q.to(torch.device("cuda", dev))
            D_loc, I_loc = self.index.search_device(q, top_docs)
            # local row -> passage id on the device; padding stays -1 so that the merge ignores it
            ids_map = self._local_ids_on(I_loc.device)
            if ids_map.numel():
                I_glob = torch.where(I_loc >= 0, ids_map[I_loc.clamp(min=0)], torch.full_like(I_loc, -1))
            else:
                I_glob = torch.full_like(I_loc, -1)
            D, I = self._sh.exchange_merge(D_loc.contiguous(), I_glob.contiguous(), lists_sorted=self._ids_increasing)
            scores, ids = D.cpu().numpy(), I.cpu().numpy()
            db_ids = ids.astype(str).tolist()
            result.extend([(db_ids[i], scores[i]) for i in range(len(db_ids))])
        return result

    def serialize(self, dir_path):
        """Every rank writes its own pair `index.{rank}of{world}.faiss` / `index_meta.{rank}of{world}.faiss`."""
        tag = f"{self.rank}of{self.world_size}"
        faiss_api.write_index(self.index, os.path.join(dir_path, f"index.{tag}.faiss"))
        with open(os.path.join(dir_path, f"index_meta.{tag}.faiss"), mode='wb') as f:
            pickle.dump(self.index_id_to_db_id, f)

    def deserialize_from(self, dir_path):
        tag = f"{self.rank}of{self.world_size}"
        self.index = faiss_api.read_index(os.path.join(dir_path, f"index.{tag}.faiss"), faiss_api.IO_FLAG_MMAP,
                                          device=getattr(self.index, "device", None))
        self._sh.index = self.index
        with open(os.path.join(dir_path, f"index_meta.{tag}.faiss"), "rb") as reader:
            self.index_id_to_db_id = pickle.load(reader)
        self._ids_dev = None
        assert len(self.index_id_to_db_id) == self.index.ntotal, \
            'Deserialized index_id_to_db_id should match faiss index size'

    def close(self):
        self._sh.close()
