"""Stages the UNMODIFIED reference modules of the hot path's callers for the C1 acceptance test.  TEST INFRASTRUCTURE ONLY.

The reference is Python, so "building oracle/_ref" (task statement, tier framing (3)) means making its own files
available where the GPU box can see them: /root/reference exists only in the build container, the `-m gpu` tests
run on a box that receives a snapshot of this repository.  This script packs the files listed below, byte for
byte, from /root/reference into ONE archive, oracle/_ref/reference_py.tar, next to a sha256 manifest.  oracle/_ref/
is git-ignored (never committed: no reference source enters the history) but travels with the snapshot, like the
built .so files.  Only tests/test_gpu_reference_files.py reads the archive: it unpacks it into a temporary
directory to run the reference's own faiss_store.py / rag_datastore_manager.py / 2-cli-rag-search.py on top of
the `faiss` shim.

    python oracle/stage_reference.py            # called by __graft_entry__.build() when /root/reference exists
"""
from __future__ import annotations

import hashlib
import json
import os
import shutil

HERE = os.path.dirname(os.path.abspath(__file__))
REFERENCE = "/root/reference"
ARCHIVE = os.path.join(HERE, "_ref", "reference_py.tar")
FILES = [
    "faiss_store.py",              # the wrapper type
    "rag_datastore_manager.py",    # RAGDatabaseManager.load_indices / search_similar_documents (a3, a4, a6)
    "database.py",                 # Database -> FAISSVectorStore() singleton user
    "2-cli-rag-search.py",         # BASELINE configs[0]: the CLI over the repo's own index
    "data/faiss_index.bin",
    "data/faiss_index.bin.mapping",
    "data/documents.db",
    "data/documents.json",
]


def stage(force: bool = False) -> str | None:
    """Packs the files; returns the archive path, or None when neither the reference tree nor an archive exists."""
    import tarfile

    if not os.path.isdir(REFERENCE):
        return ARCHIVE if os.path.exists(ARCHIVE) else None
    os.makedirs(os.path.dirname(ARCHIVE), exist_ok=True)
    manifest = {rel: hashlib.sha256(open(os.path.join(REFERENCE, rel), "rb").read()).hexdigest() for rel in FILES}
    meta = os.path.join(os.path.dirname(ARCHIVE), "MANIFEST.json")
    if not force and os.path.exists(ARCHIVE) and os.path.exists(meta):
        try:
            if json.load(open(meta)).get("sha256") == manifest:
                return ARCHIVE
        except Exception:
            pass
    with tarfile.open(ARCHIVE, "w") as tar:
        for rel in FILES:
            tar.add(os.path.join(REFERENCE, rel), arcname=rel)
    with open(meta, "w") as fh:
        json.dump({"source": REFERENCE, "sha256": manifest}, fh, indent=1)
    return ARCHIVE


def unpack(dst: str) -> str | None:
    """The reference's files under `dst` (a scratch directory): copied from the live tree in the build container,
    else unpacked from the staged archive; None when neither exists (the test then skips)."""
    import tarfile

    if os.path.isdir(REFERENCE):
        for rel in FILES:
            out = os.path.join(dst, rel)
            os.makedirs(os.path.dirname(out), exist_ok=True)
            shutil.copyfile(os.path.join(REFERENCE, rel), out)
        return dst
    if os.path.exists(ARCHIVE):
        with tarfile.open(ARCHIVE) as tar:
            tar.extractall(dst, filter="data")
        return dst
    return None


if __name__ == "__main__":
    print(stage(force=True))
