#!/usr/bin/env python
"""BASELINE configs[0]: the reference's own, unmodified 2-cli-rag-search.py / rag_datastore_manager.py /
faiss_store.py over the repo's own data/faiss_index.bin on the `faiss` shim; prints one JSON line with the top-5
of the known-answer query and the per-query latency (35 kB index: latency only, no roofline meaning).

    python tools/c1_reference_cli.py > gpurun_out/c1_cli.json
"""
import json
import os
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    from oracle import stage_reference

    with tempfile.TemporaryDirectory() as work:
        if stage_reference.unpack(work) is None:
            print(json.dumps({"c1": "unavailable", "why": "no /root/reference and no oracle/_ref/reference_py.tar"}))
            return
        env = dict(os.environ)
        shim = os.path.join(ROOT, "rag-faiss-embedding_b200", "shim")
        env["PYTHONPATH"] = os.pathsep.join([shim, ROOT] + ([env["PYTHONPATH"]] if env.get("PYTHONPATH") else []))
        r = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "ref_driver.py"),
                            os.path.join(ROOT, "tests", "golden", "fixture_answers.json")],
                           cwd=work, env=env, capture_output=True, text=True, timeout=900)
        if r.returncode != 0:
            print(json.dumps({"c1": "failed", "stderr": r.stderr[-2000:]}))
            sys.exit(1)
        o = json.loads([ln for ln in r.stdout.splitlines() if ln.startswith("REF_DRIVER_JSON ")][-1][16:])
    print(json.dumps({
        "config": "repo's own data/faiss_index.bin (23 x 384 fp32, IndexFlatL2), top-5 via the unmodified "
                  "2-cli-rag-search.py -> rag_datastore_manager.py -> `import faiss` (shim) -> libb200flat.so",
        "query": "row:0 (the index's own row 0; the sentence encoder is stubbed, no weights offline)",
        "top5_doc_ids": o["rag"]["row:0"]["ids"], "top5_sq_distances": o["rag"]["row:0"]["dist"],
        "latency_ms_per_query_median": o["cli_latency_ms_median"],
        "latency_includes": "stub embedding + index.search(nq=1, k=5) on the B200 (host buffers, synchronous) + "
                            "mapping unpickle + 5 sqlite fetches",
        "index_file_rewritten_byte_identical": bool(o["resaved_identical"] and o["manager_saved_identical"]),
    }))


if __name__ == "__main__":
    main()
