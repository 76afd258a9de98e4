"""The HTTP surface the reference's launcher expects but never defines.

`3-fastapi-uvicorn-server.sh:49-56` runs `uvicorn.run("query:app")`, `4-api-rag-search.py:91-107` posts
`{"text", "top_k"}` to `/search` and reads `{"similar_documents": [{title, url, content, score}],
"generated_response"}`, `health.sh:3` curls `/health` -- yet `query.py` has no `app`.  This module builds
that app around any vector store with the FAISSVectorStore surface.  Requests arriving within a short
window are micro-batched into ONE index.search call, which moves a busy server from the nq = 1 scan to
the tcgen05 path (nq >= 2).

    # query.py (one added line makes the reference's launcher work)
    from rag_faiss_embedding_b200.server import create_app
    app = create_app(embed=my_encoder, store=my_store, fetch_document=my_db.get_document_by_id)
"""
import asyncio
from typing import Awaitable, Callable, Dict, List, Optional, Sequence

import numpy as np


try:  # the request model lives at module level so FastAPI can resolve it as the JSON body
    from pydantic import BaseModel

    class SearchRequest(BaseModel):
        text: str
        top_k: int = 5
except ImportError:  # pragma: no cover - the server is optional
    SearchRequest = None


class MicroBatcher:
    """Collects single-query searches for up to `max_wait_ms` (or `max_batch`) and runs them as one batch."""

    def __init__(self, search_many: Callable[[np.ndarray, int], tuple], max_batch: int = 128, max_wait_ms: float = 2.0):
        self._search_many = search_many
        self.max_batch = max_batch
        self.max_wait = max_wait_ms / 1e3
        self._pending: List[tuple] = []
        self._lock = asyncio.Lock()
        self._flusher: Optional[asyncio.Task] = None
        self.batches_run = 0
        self.largest_batch = 0

    async def search(self, vector: np.ndarray, k: int):
        loop = asyncio.get_running_loop()
        fut: asyncio.Future = loop.create_future()
        async with self._lock:
            self._pending.append((np.asarray(vector, np.float32).reshape(-1), int(k), fut))
            if len(self._pending) >= self.max_batch:
                await self._flush_locked()
            elif self._flusher is None or self._flusher.done():
                self._flusher = loop.create_task(self._flush_later())
        return await fut

    async def _flush_later(self):
        await asyncio.sleep(self.max_wait)
        async with self._lock:
            await self._flush_locked()

    async def _flush_locked(self):
        if not self._pending:
            return
        batch, self._pending = self._pending, []
        kmax = max(k for _, k, _ in batch)
        q = np.stack([v for v, _, _ in batch])
        try:
            dist, ids = await asyncio.get_running_loop().run_in_executor(None, self._search_many, q, kmax)
        except Exception as exc:  # noqa: BLE001 - every waiter must be released
            for _, _, fut in batch:
                if not fut.done():
                    fut.set_exception(exc)
            return
        self.batches_run += 1
        self.largest_batch = max(self.largest_batch, len(batch))
        for row, (_, k, fut) in enumerate(batch):
            if not fut.done():
                fut.set_result((np.asarray(dist[row])[:k], list(ids[row])[:k]))


def create_app(embed: Callable[[Sequence[str]], np.ndarray], store, fetch_document: Callable[[int], Optional[Dict]],
               generate: Optional[Callable[[str, List[Dict]], Awaitable[str]]] = None, max_batch: int = 128,
               max_wait_ms: float = 2.0):
    """FastAPI app with the request / response shapes of 4-api-rag-search.py.

    embed(texts) -> [n, d] float32; store has search_many(queries, k) -> (D [n, k], ids per query);
    fetch_document(doc_id) -> {"title", "url", "content", ...} or None; score = 1 / (1 + distance), the
    reference's formula (query.py:42, 2-cli-rag-search.py:48).
    """
    from fastapi import FastAPI

    app = FastAPI(title="rag-faiss-embedding on B200")
    batcher = MicroBatcher(store.search_many, max_batch=max_batch, max_wait_ms=max_wait_ms)
    app.state.batcher = batcher

    @app.get("/health")
    async def health():
        index = getattr(store, "index", None)
        return {"status": "ok", "vectors": int(getattr(index, "ntotal", 0)), "batches_run": batcher.batches_run,
                "largest_batch": batcher.largest_batch}

    @app.post("/search")
    async def search(req: SearchRequest):
        vec = np.asarray(embed([req.text]), np.float32)[0]
        dist, ids = await batcher.search(vec, max(1, req.top_k))
        docs = []
        for d, doc_id in zip(dist, ids):
            doc = fetch_document(int(doc_id))
            if doc:
                docs.append({"title": doc.get("title", ""), "url": doc.get("url", ""), "content": doc.get("content", ""),
                             "score": float(1.0 / (1.0 + float(d)))})
        answer = ""
        if generate is not None and docs:
            answer = await generate(req.text, docs)
        return {"similar_documents": docs, "generated_response": answer}

    return app
