"""Runs INSIDE a scratch copy of the reference tree (cwd), with the `faiss` shim first on PYTHONPATH: drives the
reference's own, unmodified faiss_store.py / database.py / rag_datastore_manager.py / 2-cli-rag-search.py and prints
what they returned as one JSON line.  Started by tests/test_gpu_reference_files.py; not a test module itself.

The only stand-in is the sentence encoder (no all-MiniLM-L6-v2 weights offline, and the encoder is outside the hot
path): `EmbeddingModel` is replaced by a stub that maps the text "row:<i>" to row i of the index and "pert:<j>" to
the j-th perturbed golden query, so the answers are the committed float64 golden answers.
"""
import hashlib
import importlib.util
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.getcwd())   # the reference's modules (database.py imports faiss_store by name)

import faiss  # noqa: E402  -- must resolve to the shim

out = {"faiss_file": faiss.__file__}
golden = json.load(open(sys.argv[1]))
perturbed = np.asarray(golden["perturbed_queries"], np.float32)
INDEX = "data/faiss_index.bin"
orig_bytes = open(INDEX, "rb").read()
out["index_sha256"] = hashlib.sha256(orig_bytes).hexdigest()

# ---- faiss_store.py, unmodified: singleton, load on construction, search, add, save, load, reset ------------------
import faiss_store  # noqa: E402

store = faiss_store.FAISSVectorStore()            # dimension=384, index_path="data/faiss_index.bin": loads it
out["store_singleton"] = store is faiss_store.FAISSVectorStore(dimension=7, index_path="nowhere")
out["store_ntotal"] = int(store.index.ntotal)
out["store_doc_ids"] = [int(i) for i in store.doc_ids]
rows = np.stack([store.index.reconstruct(i) for i in range(store.index.ntotal)])
out["store_search"] = {}
for r in (0, 1, 2, 7, 22):
    dist, ids = store.search(rows[r], k=5)        # ndarray query
    out["store_search"][str(r)] = {"dist": [float(x) for x in dist], "ids": [int(i) for i in ids]}
dist, ids = store.search(list(map(float, rows[3])), k=40)   # list query, k > ntotal: -1 rows are dropped
out["store_k40"] = {"n": len(ids), "n_dist": int(len(dist))}
os.makedirs("out", exist_ok=True)
store.save_index("out/resaved.bin")               # faiss.write_index + pickle of the mapping
out["resaved_identical"] = open("out/resaved.bin", "rb").read() == orig_bytes
out["mapping_identical"] = open("out/resaved.bin.mapping", "rb").read() == open(INDEX + ".mapping", "rb").read()
new_rows = (rows[:2] + 0.125).astype(np.float32)
store.add_vectors([list(map(float, v)) for v in new_rows], [101, 102])   # list-of-lists path (np.array inside)
store.add_vectors(rows[5] - 0.25, [103])                                  # 1-D path (reshape inside)
out["after_add_ntotal"] = int(store.index.ntotal)
dist, ids = store.search(new_rows[1], k=3)
out["after_add_top"] = {"ids": [int(i) for i in ids], "d0": float(dist[0])}
store.save_index("out/grown.bin")
store.load_index("out/grown.bin")
out["reloaded_ntotal"] = int(store.index.ntotal)
out["reloaded_doc_ids_tail"] = [int(i) for i in store.doc_ids[-3:]]
dist, ids = store.search(rows[5] - 0.25, k=1)
out["reloaded_top"] = {"ids": [int(i) for i in ids], "d0": float(dist[0])}
store.reset()
out["reset_ntotal"] = int(store.index.ntotal)
dist, ids = store.search(rows[0], k=5)            # empty index: every label is -1 -> nothing survives
out["reset_search_len"] = len(ids)
store.load_index(INDEX)                           # back to the shipped index for database.py below

# ---- database.py, unmodified: Database() builds its FAISSVectorStore() (the same singleton) ------------------------
import database  # noqa: E402

db = database.Database()
out["db_store_is_singleton"] = db.vector_store is store
dist, ids = db.vector_store.search(rows[22], 5)
out["db_search_22"] = {"dist": [float(x) for x in dist], "ids": [int(i) for i in ids]}
out["db_doc"] = (db.get_document_by_id(int(ids[0])) or {}).get("id")

# ---- rag_datastore_manager.py + 2-cli-rag-search.py, unmodified (BASELINE configs[0]) ----------------------------
import rag_datastore_manager as rdm  # noqa: E402


class StubEmbeddingModel:
    """Stand-in for the sentence encoder only (the reference's class loads all-MiniLM-L6-v2 from the hub)."""

    def generate_embeddings(self, texts, batch_size=32):
        vecs = []
        for t in texts:
            kind, idx = t.split(":")
            vecs.append(rows[int(idx)] if kind == "row" else perturbed[int(idx)])
        return np.array(vecs)


rdm.EmbeddingModel = StubEmbeddingModel
spec = importlib.util.spec_from_file_location("cli_rag_search", "2-cli-rag-search.py")
cli_mod = importlib.util.module_from_spec(spec)
spec.loader.exec_module(cli_mod)
cli = cli_mod.CLISearch()                         # RAGDatabaseManager() + load_indices(): faiss.read_index
out["cli_index_ntotal"] = int(cli.rag_manager.faiss_index.ntotal)
out["rag"] = {}
import asyncio  # noqa: E402
import time  # noqa: E402

lat = []
for text in ("row:0", "row:7", "pert:0", "pert:5"):
    t0 = time.perf_counter()
    res = asyncio.run(cli.search(text))           # search_similar_documents: search + mapping + sqlite fetch
    lat.append(time.perf_counter() - t0)
    out["rag"][text] = {"ids": [int(d["id"]) for d in res], "dist": [float(d["distance"]) for d in res],
                        "titles": [d["title"] for d in res]}
out["cli_latency_ms_median"] = round(sorted(lat)[len(lat) // 2] * 1e3, 3)
import builtins  # noqa: E402

builtins.input = lambda *a, **k: ""               # print_results asks for a document number
cli.print_results(asyncio.run(cli.search("row:0")))
mgr = rdm.RAGDatabaseManager()
mgr.load_indices()
res = mgr.search_similar_documents("row:2", k=3)
out["rag_k3"] = {"ids": [int(d["id"]) for d in res], "dist": [float(d["distance"]) for d in res]}
# _save_faiss_index: write_index + mapping, from the manager's own index
docs = json.load(open("data/documents.json"))
os.replace(INDEX, "out/orig.bin")
mgr._save_faiss_index([{"id": i} for i in golden["mapping"]])
out["manager_saved_identical"] = open(INDEX, "rb").read() == orig_bytes
out["n_documents_json"] = len(docs)
cli.cleanup()
print("REF_DRIVER_JSON " + json.dumps(out))
