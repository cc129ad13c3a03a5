"""`Indexer` with the reference's interface, on the B200 library.

Mirror of /root/reference/retriever/index.py:17-83 (same method names,
argument meaning, return types and error behaviour) for callers that import
an indexer class directly instead of going through `kirag_b200.as_faiss`.
Differences are confined to host-side glue that the reference does in Python
loops: ids are mapped with one vectorised numpy take instead of an n*k
`str()` list comprehension per element (index.py:49), and the id map grows by
chunks instead of `np.concatenate` per call (index.py:81-83).
"""
from __future__ import annotations

import logging
import os
import pickle
from typing import List, Tuple

import numpy as np

from . import faiss_api

logger = logging.getLogger()

FAISSINDEX_DICT = {
    "inner_product": faiss_api.IndexFlatIP,
    "l2": faiss_api.IndexFlatL2,
}


class Indexer(object):

    def __init__(self, vector_sz, metric="inner_product", n_subquantizers=0, n_bits=8, device=None):
        if n_subquantizers > 0:
            self.index = faiss_api.IndexPQ(vector_sz, n_subquantizers, n_bits, faiss_api.METRIC_INNER_PRODUCT)
        else:
            self.index = FAISSINDEX_DICT[metric](vector_sz) if device is None else \
                FAISSINDEX_DICT[metric](vector_sz, device=device)
        self.index_id_to_db_id = np.empty((0), dtype=np.int64)

    def index_data(self, ids, embeddings):
        self._update_id_mapping(ids)
        embeddings = embeddings.astype('float32')
        if not self.index.is_trained:
            self.index.train(embeddings)
        self.index.add(embeddings)
        logger.info(f'Total data indexed {len(self.index_id_to_db_id)}')

    def search_knn(self, query_vectors: np.array, top_docs: int, index_batch_size=1024,
                   verbose: bool = True) -> List[Tuple[List[object], List[float]]]:
        query_vectors = query_vectors.astype('float32')
        result = []
        nbatch = (len(query_vectors) - 1) // index_batch_size + 1
        for k in range(nbatch):
            start_idx = k * index_batch_size
            end_idx = min((k + 1) * index_batch_size, len(query_vectors))
            q = query_vectors[start_idx:end_idx]
            scores, indexes = self.index.search(q, top_docs)
            # convert to external ids; -1 padding indexes the LAST id exactly like
            # the reference's index_id_to_db_id[-1] does (index.py:49)
            db_ids = self.index_id_to_db_id[indexes].astype(str).tolist() if len(self.index_id_to_db_id) else \
                [[str(i) for i in row] for row in indexes]
            result.extend([(db_ids[i], scores[i]) for i in range(len(db_ids))])
        return result

    def serialize(self, dir_path):
        index_file = os.path.join(dir_path, "index.faiss")
        meta_file = os.path.join(dir_path, "index_meta.faiss")
        logger.info(f'Serializing index to {index_file}, meta data to {meta_file}')
        faiss_api.write_index(self.index, index_file)
        with open(meta_file, mode='wb') as f:
            pickle.dump(self.index_id_to_db_id, f)

    def deserialize_from(self, dir_path):
        index_file = os.path.join(dir_path, "index.faiss")
        meta_file = os.path.join(dir_path, "index_meta.faiss")
        logger.info(f'Loading index from {index_file}, meta data from {meta_file}')
        self.index = faiss_api.read_index(index_file, faiss_api.IO_FLAG_MMAP)
        logger.info('Loaded index of type %s and size %d', type(self.index), self.index.ntotal)
        with open(meta_file, "rb") as reader:
            self.index_id_to_db_id = pickle.load(reader)
        assert len(
            self.index_id_to_db_id) == self.index.ntotal, 'Deserialized index_id_to_db_id should match faiss index size'

    def _update_id_mapping(self, db_ids: List):
        new_ids = np.array(db_ids, dtype=np.int64)
        self.index_id_to_db_id = np.concatenate((self.index_id_to_db_id, new_ids), axis=0)


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
