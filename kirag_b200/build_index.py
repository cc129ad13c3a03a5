"""Build an index from saved corpus embeddings.

Mirror of /root/reference/faiss_index_corpus.py:23-52 (`build_faiss_index`):
glob `corpus_embeddings_{s}_{e}.pkl` / `passage_id_list_{s}_{e}.pkl` written by
compute_corpus_embeddings.py:114-115, add them shard by shard in ascending
order of the end index, serialise `index.faiss` + `index_meta.faiss`, delete
the pickles.

Two reference quirks are fixed by construction, not reproduced:
  * it pairs files by SUBSTRING (faiss_index_corpus.py:37-41: end id "999999"
    also matches passage_id_list_1000000_1999999.pkl) — here pairs are matched
    on the exact `{start}_{end}` suffix;
  * it parses the end index with split(".")[0] (:24), which breaks when a
    directory name contains a dot — here only the basename is parsed.
"""
from __future__ import annotations

import glob
import logging
import os
import pickle
import re
from typing import List, Tuple

from .index import Indexer

logger = logging.getLogger(__file__)

_EMB_RE = re.compile(r"^corpus_embeddings_(\d+)_(\d+)\.pkl$")
_IDS_RE = re.compile(r"^passage_id_list_(\d+)_(\d+)\.pkl$")


def pair_embedding_files(index_folder: str) -> List[Tuple[str, str]]:
    emb = {}
    for f in glob.glob(os.path.join(index_folder, "corpus_embeddings_*.pkl")):
        m = _EMB_RE.match(os.path.basename(f))
        if m:
            emb[(int(m.group(1)), int(m.group(2)))] = f
    ids = {}
    for f in glob.glob(os.path.join(index_folder, "passage_id_list_*.pkl")):
        m = _IDS_RE.match(os.path.basename(f))
        if m:
            ids[(int(m.group(1)), int(m.group(2)))] = f
    assert len(emb) == len(ids), "embedding / passage-id file counts differ"
    pairs = []
    for key in sorted(emb, key=lambda se: se[1]):
        assert key in ids, f"no passage_id_list file for corpus_embeddings_{key[0]}_{key[1]}.pkl"
        pairs.append((emb[key], ids[key]))
    return pairs


def build_faiss_index(args=None, index_folder: str = None, embedding_size: int = 1024,
                      delete_inputs: bool = True, device=None) -> Indexer:
    """Same entry point name as the reference; `args` may be its argparse namespace."""
    if args is not None:
        index_folder = args.index_folder
        embedding_size = args.embedding_size
    indexer = Indexer(embedding_size, metric="inner_product", device=device)
    pairs = pair_embedding_files(index_folder)
    total = 0
    for emb_file, _ in pairs:
        m = _EMB_RE.match(os.path.basename(emb_file))
        total += int(m.group(2)) - int(m.group(1)) + 1
    if total > 0:
        indexer.index.reserve(total)  # one allocation for the whole corpus
    for emb_file, ids_file in pairs:
        with open(emb_file, "rb") as f:
            embeddings = pickle.load(f)
        with open(ids_file, "rb") as f:
            passage_ids = pickle.load(f)
        indexer.index_data(passage_ids, embeddings.cpu().numpy())
    logger.info(f"Saving index to {index_folder} ... ")
    indexer.serialize(index_folder)
    if delete_inputs:
        for emb_file, ids_file in pairs:
            os.remove(emb_file)
            os.remove(ids_file)
    return indexer
