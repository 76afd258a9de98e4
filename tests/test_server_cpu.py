"""CPU tier: the HTTP app the reference's launcher expects (3-fastapi-uvicorn-server.sh, 4-api-rag-search.py,
health.sh) -- request/response shapes, the reference's score formula, and micro-batching of concurrent
requests into one batched store call.  The store is the FAISSVectorStore mirror over an oracle-backed index
(host logic only)."""
import asyncio

import numpy as np
import pytest

import oracle as orc
from tests.helpers import OracleIndex


@pytest.fixture()
def app_and_store(monkeypatch, tmp_path):
    from rag_faiss_embedding_b200 import store as st
    from rag_faiss_embedding_b200.server import create_app

    monkeypatch.setattr(st, "IndexFlatL2", OracleIndex)
    s = st.FAISSVectorStore(dimension=16, index_path=str(tmp_path / "none.bin"))
    x = orc.np_synth_rows(5, 0, 40, 16)
    s.add_vectors(x, [100 + i for i in range(40)])
    docs = {100 + i: {"id": 100 + i, "title": f"t{i}", "url": f"u{i}", "content": f"c{i}"} for i in range(40)}
    calls = []
    orig = s.search_many

    def counting(q, k):
        calls.append(len(q))
        return orig(q, k)

    s.search_many = counting

    def embed(texts):   # "text" is the row number of the vector to look for
        return np.stack([x[int(t)] for t in texts])

    app = create_app(embed, s, docs.get, max_batch=64, max_wait_ms=20.0)
    return app, x, calls


def test_search_and_health_shapes(app_and_store):
    from fastapi.testclient import TestClient

    app, x, calls = app_and_store
    with TestClient(app) as client:
        r = client.post("/search", json={"text": "7", "top_k": 3})
        assert r.status_code == 200
        body = r.json()
        assert set(body) == {"similar_documents", "generated_response"}
        docs = body["similar_documents"]
        assert len(docs) == 3 and docs[0]["title"] == "t7" and docs[0]["score"] == 1.0   # 1 / (1 + 0)
        assert all(set(d) == {"title", "url", "content", "score"} for d in docs)
        assert docs[0]["score"] >= docs[1]["score"] >= docs[2]["score"]
        h = client.get("/health").json()
        assert h["status"] == "ok" and h["vectors"] == 40 and h["batches_run"] >= 1


def test_concurrent_requests_are_micro_batched(app_and_store):
    import httpx

    app, x, calls = app_and_store

    async def run():
        transport = httpx.ASGITransport(app=app)
        async with httpx.AsyncClient(transport=transport, base_url="http://test") as client:
            rs = await asyncio.gather(*[client.post("/search", json={"text": str(i), "top_k": 2}) for i in range(12)])
        return [r.json() for r in rs]

    out = asyncio.run(run())
    assert [o["similar_documents"][0]["title"] for o in out] == [f"t{i}" for i in range(12)]
    assert max(calls) > 1, f"requests were not batched: {calls}"
    assert sum(calls) == 12
