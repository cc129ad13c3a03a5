"""faiss.write_index / faiss.read_index (retriever/index.py:62,73) at the size SURVEY §8 a5 states: the 21M x 1024 fp32
index (86 GB `index.faiss`).  Needs ~90 GB of scratch disk; prints one JSON line.  python tools/index_io_21m.py [ROWS]"""
import json
import os
import shutil
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import bench  # noqa: E402
from kirag_b200 import faiss_api  # noqa: E402

rows = int(sys.argv[1]) if len(sys.argv) > 1 else 21_000_000
dev = torch.device("cuda", 0)
tmp = tempfile.mkdtemp(prefix="kirag_io_", dir=os.environ.get("KIRAG_IO_DIR"))
try:
    free_disk = shutil.disk_usage(tmp).free
    need = rows * 1024 * 4
    if need > free_disk * 0.9:
        print(json.dumps({"skipped": f"{free_disk >> 30} GiB free on {tmp}, need {need >> 30}"}))
        sys.exit(0)
    ix = faiss_api.IndexFlatIP(1024, device=0)
    t0 = time.perf_counter()
    bench.build_shard(ix, 0, rows, dev)  # no reserve: grows in place like the reference's build loop
    build_s = time.perf_counter() - t0
    g = torch.Generator(device=dev)
    g.manual_seed(4321)
    q = torch.nn.functional.normalize(torch.randn(8, 1024, generator=g, device=dev), dim=1)
    D0, I0 = ix.search_device(q, 100)
    path = os.path.join(tmp, "index.faiss")
    t0 = time.perf_counter()
    faiss_api.write_index(ix, path)
    write_s = time.perf_counter() - t0
    size = os.path.getsize(path)
    del ix
    torch.cuda.empty_cache()
    t0 = time.perf_counter()
    ix2 = faiss_api.read_index(path, faiss_api.IO_FLAG_MMAP, device=0)
    load_s = time.perf_counter() - t0
    D1, I1 = ix2.search_device(q, 100)
    gb = size / 1e9
    print(json.dumps({"rows": rows, "file_gb": gb, "build_s_from_device_chunks_no_reserve": build_s, "write_s": write_s,
                      "write_gbs": gb / write_s, "load_s": load_s, "load_gbs": gb / load_s,
                      "ntotal_after_load": ix2.ntotal,
                      "same_top100_after_reload": bool(torch.equal(I0, I1) and torch.equal(D0, D1)),
                      "note": "load = fread of the IxFI file through two pinned 64 MB buffers + one convert pass "
                              "(builds the bf16 shadow); page cache as left by the write"}))
finally:
    shutil.rmtree(tmp, ignore_errors=True)
