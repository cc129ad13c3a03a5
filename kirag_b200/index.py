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
