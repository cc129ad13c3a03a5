"""Snapshot the UNMODIFIED reference files of the hot path into baseline/_ref/ (build container only).

`/root/reference` does not exist on the GPU box, so the tests that run the reference's own callers
(retriever/index.py, retriever/retrievers.py, compute_corpus_embeddings.py, faiss_index_corpus.py) on
top of `kirag_b200.as_faiss` could only ever skip there.  `baseline/_ref/` is the base contract's
place for the unmodified reference: git-ignored (never in history), NOT gpurun-ignored (it travels to
the GPU box like the built .so files).  The reference has no setup.py / pyproject.toml, so
`pip install --target baseline/_ref /root/reference` fails ("Neither 'setup.py' nor 'pyproject.toml'
found"); a plain byte-for-byte copy of the modules the hot path imports is the install.

A manifest with the sha256 of every copied file is written next to them; the tests check the copies
against it, so a stale or edited snapshot is noticed.
"""
from __future__ import annotations

import hashlib
import json
import os
import shutil

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REFERENCE = "/root/reference"
DEST = os.path.join(ROOT, "baseline", "_ref")
# everything `retriever.retrievers`, `compute_corpus_embeddings` and `faiss_index_corpus` import
PATHS = ["retriever", "dataset", "utils", "compute_corpus_embeddings.py", "faiss_index_corpus.py", "requirements.txt"]


def _sha256(path: str) -> str:
    h = hashlib.sha256()
    with open(path, "rb") as f:
        for block in iter(lambda: f.read(1 << 20), b""):
            h.update(block)
    return h.hexdigest()


def snapshot(force: bool = False) -> str | None:
    """Copy PATHS from /root/reference to baseline/_ref.  Returns DEST, or None where there is no reference."""
    if not os.path.isdir(os.path.join(REFERENCE, "retriever")):
        return DEST if os.path.isdir(os.path.join(DEST, "retriever")) else None
    manifest = {}
    os.makedirs(DEST, exist_ok=True)
    for rel in PATHS:
        src = os.path.join(REFERENCE, rel)
        if os.path.isdir(src):
            for dirpath, dirnames, filenames in os.walk(src):
                dirnames[:] = [d for d in dirnames if d != "__pycache__"]
                for fn in filenames:
                    if fn.endswith(".pyc"):
                        continue
                    s = os.path.join(dirpath, fn)
                    r = os.path.relpath(s, REFERENCE)
                    manifest[r] = _sha256(s)
        elif os.path.isfile(src):
            manifest[rel] = _sha256(src)
    for r, digest in manifest.items():
        dst = os.path.join(DEST, r)
        if not force and os.path.isfile(dst) and _sha256(dst) == digest:
            continue
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        shutil.copyfile(os.path.join(REFERENCE, r), dst)
    with open(os.path.join(DEST, "SNAPSHOT.json"), "w") as f:
        json.dump({"source": REFERENCE, "files": manifest}, f, indent=1, sort_keys=True)
    return DEST


def verify(dest: str = DEST) -> bool:
    """True if every file listed in the manifest is present with the recorded digest."""
    path = os.path.join(dest, "SNAPSHOT.json")
    if not os.path.isfile(path):
        return False
    files = json.load(open(path))["files"]
    return all(os.path.isfile(os.path.join(dest, r)) and _sha256(os.path.join(dest, r)) == d for r, d in files.items())


if __name__ == "__main__":
    print(snapshot(), "verified" if verify() else "NOT verified")
