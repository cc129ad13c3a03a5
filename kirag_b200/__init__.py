"""kirag_b200 — B200-native (sm_100a) exact inner-product top-k search and
mean-pool + L2-normalise epilogue behind the interfaces KiRAG takes from
`faiss` and from `retriever/encoders.py`.

    import kirag_b200.as_faiss          # `import faiss` now resolves to this library
    from kirag_b200 import Indexer      # or use the Indexer mirror directly
    from kirag_b200.pooling import average_pool, e5_embed, bge_embed

Importing this package loads no CUDA code; the first use of an index or of
the pooling op loads `libkirag_b200.so` and fails loudly if it is missing.
"""
from . import _lib  # noqa: F401
from .faiss_api import IndexFlatIP, read_index, write_index  # noqa: F401
from .index import Indexer, ShardedIndexer  # noqa: F401

__all__ = ["IndexFlatIP", "Indexer", "ShardedIndexer", "read_index", "write_index"]
__version__ = "0.1.0"
