"""GPU tier: BASELINE configs[0] / SURVEY 8 (a8, d-C1) acceptance -- the reference's OWN files, unmodified, running on
the `faiss` shim (rag-faiss-embedding_b200/shim) with the engine underneath:

  faiss_store.py           FAISSVectorStore singleton: construct (read_index), search, add_vectors, save_index,
                           load_index, reset                                (faiss_store.py:4,29,46,64,91,106,126)
  database.py              Database() -> FAISSVectorStore()
  rag_datastore_manager.py RAGDatabaseManager.load_indices / search_similar_documents / _save_faiss_index
                                                                     (rag_datastore_manager.py:8,138,186,205,218)
  2-cli-rag-search.py      CLISearch over the repo's own data/faiss_index.bin, top-5

The files come from /root/reference when it exists (build container) or from the archive that
oracle/stage_reference.py packed (git-ignored, travels with the snapshot); without either the test skips.  They run
in a subprocess (tests/ref_driver.py) inside a scratch copy of the tree, so their singletons and relative paths
behave as in the reference checkout.  Expected answers: tests/golden/fixture_answers.json (float64 brute force over
the reference's FAISS-written index)."""
import json
import os
import subprocess
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def driver_output(tmp_path_factory):
    from oracle import stage_reference

    work = tmp_path_factory.mktemp("reference_checkout")
    if stage_reference.unpack(str(work)) is None:
        pytest.skip("neither /root/reference nor oracle/_ref/reference_py.tar is present")
    env = dict(os.environ)
    shim = os.path.join(ROOT, "rag-faiss-embedding_b200", "shim")
    env["PYTHONPATH"] = os.pathsep.join([shim, ROOT] + ([env["PYTHONPATH"]] if env.get("PYTHONPATH") else []))
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "ref_driver.py"),
                        os.path.join(ROOT, "tests", "golden", "fixture_answers.json")],
                       cwd=str(work), env=env, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-3000:] + "\n" + r.stderr[-3000:]
    line = [ln for ln in r.stdout.splitlines() if ln.startswith("REF_DRIVER_JSON ")][-1]
    return json.loads(line[len("REF_DRIVER_JSON "):])


def _case(golden, metric, k, queries):
    return next(c for c in golden["cases"] if c["metric"] == metric and c["k"] == k and c["queries"] == queries)


def test_reference_faiss_store_runs_unmodified_on_the_shim(driver_output, golden):
    o = driver_output
    assert "rag-faiss-embedding_b200" in o["faiss_file"]          # `import faiss` resolved to the shim
    assert o["index_sha256"] == golden["sha256_index"]
    assert o["store_singleton"] is True                           # faiss_store.py:14-22
    assert o["store_ntotal"] == 23 and o["store_doc_ids"] == golden["mapping"]
    case = _case(golden, 1, 5, "self")
    for r, got in o["store_search"].items():
        r = int(r)
        assert got["ids"] == [golden["mapping"][i] for i in case["ids"][r]], r
        assert np.allclose(got["dist"], case["dists"][r], rtol=1e-5, atol=1e-4), r
    assert o["store_search"]["0"]["ids"] == [9, 11, 14, 21, 8]    # SURVEY 8c known answer
    assert o["store_k40"] == {"n": 23, "n_dist": 23}              # -1 padding dropped (faiss_store.py:71)
    assert o["resaved_identical"] and o["mapping_identical"]      # write_index: byte-identical to FAISS's own file
    assert o["after_add_ntotal"] == 26 and o["after_add_top"]["ids"][0] == 102 and o["after_add_top"]["d0"] == 0.0
    assert o["reloaded_ntotal"] == 26 and o["reloaded_doc_ids_tail"] == [101, 102, 103]
    assert o["reloaded_top"] == {"ids": [103], "d0": 0.0}
    assert o["reset_ntotal"] == 0 and o["reset_search_len"] == 0
    assert o["db_store_is_singleton"] is True
    c22 = case["ids"][22]
    assert o["db_search_22"]["ids"] == [golden["mapping"][i] for i in c22]
    assert o["db_doc"] == golden["mapping"][c22[0]]


def test_reference_rag_manager_and_cli_run_unmodified_on_the_shim(driver_output, golden):
    o = driver_output
    assert o["cli_index_ntotal"] == 23
    self5, pert5 = _case(golden, 1, 5, "self"), _case(golden, 1, 5, "perturbed")
    for text, got in o["rag"].items():
        kind, idx = text.split(":")
        case = self5 if kind == "row" else pert5
        want_ids = [golden["mapping"][i] for i in case["ids"][int(idx)]]
        assert got["ids"] == want_ids, text                       # row -> doc id -> sqlite document
        assert np.allclose(got["dist"], case["dists"][int(idx)], rtol=1e-5, atol=1e-4), text
        assert all(isinstance(t, str) and t for t in got["titles"])
    assert o["rag_k3"]["ids"] == [golden["mapping"][i] for i in self5["ids"][2][:3]]
    assert o["manager_saved_identical"]                           # rag_datastore_manager.py:186
    assert o["n_documents_json"] == 23
    assert o["cli_latency_ms_median"] > 0
