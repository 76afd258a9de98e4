"""GPU tier: the callers either side of the hot path over the REAL index --
FAISSVectorStore.search_many (a batch of queries in one device call) and the HTTP app the reference's launcher
expects (POST /search, GET /health) with request micro-batching, both against the oracle."""
import asyncio
import sqlite3

import numpy as np
import pytest

import oracle as orc

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def m():
    import rag_faiss_embedding_b200 as mod

    assert mod.device_count() > 0, "no B200 visible"
    return mod


def test_search_many_on_the_real_index(m, golden, tmp_path):
    s = m.FAISSVectorStore(dimension=384, index_path=golden["index_path"])
    xb, _ = orc.np_read_index(golden["index_path"])
    # the whole fixture as ONE batch: the same rows -> doc ids as 23 single-query searches and as the golden answers
    dist, ids = s.search_many(xb, k=5)
    case = next(c for c in golden["cases"] if c["metric"] == 1 and c["k"] == 5 and c["queries"] == "self")
    assert ids == [[golden["mapping"][i] for i in row] for row in case["ids"]]
    assert np.allclose(dist, np.asarray(case["dists"]), rtol=1e-5, atol=1e-4)
    for r in (0, 7, 22):
        d1, i1 = s.search(xb[r], k=5)
        assert i1 == ids[r] and np.allclose(d1, dist[r], rtol=1e-6, atol=1e-6)
    # a large batch on a larger store goes through the tensor path; k > ntotal drops the -1 rows per query
    n, d, nq, k = 30000, 384, 700, 10
    big = m.FAISSVectorStore(dimension=d, index_path=str(tmp_path / "absent.bin"))
    x = orc.c_synth_rows(1234, 0, n, d)
    big.add_vectors(x, list(range(1000, 1000 + n)))
    xq = orc.c_synth_rows(5678, 0, nq, d)
    D, mapped = big.search_many(xq, k)
    D_ref, I_ref = orc.np_search_f64(x, xq, k, 1)
    assert mapped == [[1000 + int(i) for i in row] for row in I_ref]
    assert np.allclose(D, D_ref, rtol=1e-5, atol=1e-4)
    assert big.index.stats()["last_algo"] == m.ALGO_TENSOR
    small = m.FAISSVectorStore(dimension=384, index_path=golden["index_path"])
    dist, ids = small.search_many(xb[:3], k=40)
    assert [len(r) for r in ids] == [23, 23, 23] and dist.shape == (3, 40)


def test_http_app_over_the_real_index(m, golden, tmp_path):
    """create_app(embed, store, fetch_document) with the shipped index and an sqlite document table: response
    shapes of 4-api-rag-search.py, the reference's score formula 1 / (1 + d), and micro-batching of concurrent
    requests into one index.search call on the GPU."""
    import httpx
    from fastapi.testclient import TestClient

    from rag_faiss_embedding_b200.server import create_app

    store = m.FAISSVectorStore(dimension=384, index_path=golden["index_path"])
    xb, _ = orc.np_read_index(golden["index_path"])
    conn = sqlite3.connect(str(tmp_path / "documents.db"), check_same_thread=False)
    conn.execute("CREATE TABLE documents (id INTEGER PRIMARY KEY, url TEXT, title TEXT, content TEXT)")
    conn.executemany("INSERT INTO documents VALUES (?, ?, ?, ?)",
                     [(i, f"http://x/{i}", f"title {i}", f"content {i}") for i in golden["mapping"]])
    conn.commit()

    def fetch(doc_id):
        row = conn.execute("SELECT id, url, title, content FROM documents WHERE id = ?", (doc_id,)).fetchone()
        return {"id": row[0], "url": row[1], "title": row[2], "content": row[3]} if row else None

    calls = []
    orig = store.search_many

    def counting(q, k):
        calls.append(len(q))
        return orig(q, k)

    store.search_many = counting
    app = create_app(lambda texts: np.stack([xb[int(t)] for t in texts]), store, fetch, max_batch=64, max_wait_ms=25.0)
    case = next(c for c in golden["cases"] if c["metric"] == 1 and c["k"] == 5 and c["queries"] == "self")
    with TestClient(app) as client:
        body = client.post("/search", json={"text": "0", "top_k": 5}).json()
        assert set(body) == {"similar_documents", "generated_response"}
        docs = body["similar_documents"]
        assert [d["title"] for d in docs] == [f"title {i}" for i in (9, 11, 14, 21, 8)]
        assert docs[0]["score"] == 1.0
        assert np.allclose([d["score"] for d in docs], 1.0 / (1.0 + np.asarray(case["dists"][0])), rtol=1e-5)
        h = client.get("/health").json()
        assert h["status"] == "ok" and h["vectors"] == 23

    async def burst():
        transport = httpx.ASGITransport(app=app)
        async with httpx.AsyncClient(transport=transport, base_url="http://test") as client:
            rs = await asyncio.gather(*[client.post("/search", json={"text": str(i), "top_k": 5}) for i in range(23)])
        return [r.json() for r in rs]

    calls.clear()
    out = asyncio.run(burst())
    for i, o in enumerate(out):
        want = [f"title {golden['mapping'][j]}" for j in case["ids"][i]]
        assert [d["title"] for d in o["similar_documents"]] == want, i
    assert sum(calls) == 23 and max(calls) > 1, f"requests were not batched: {calls}"
