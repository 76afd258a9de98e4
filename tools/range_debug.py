"""Diagnostics for the failure paths behind the tensor pass on near-duplicate data: four searches, stats after each.

usage: range_debug.py [metric] [layout] [d] [k] [storage: f32 | bf16] [nq]
  layout "shuffled": 300 clusters x 80 near-duplicates in random row order  -> uncertified queries WITH a k-th distance
                     (the range pass serves them from the second search on)
         "ordered" : the same rows, every cluster stored contiguously       -> thresholds never tighten, lists overflow;
                     the index switches to per-thread heaps + range pass (order-robust mode)
         "dense"   : 60 clusters x 400 near-duplicates, random order        -> the lists cut inside the cluster: the
                     extended certification fails too, the range pass lists the whole cluster
         "huge"    : 6 clusters x 4000 near-duplicates, random order        -> more list entries than the merge stages, and
                     more than the range pass lists: exact scans, the range pass switches itself off again
"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import rag_faiss_embedding_b200 as m  # noqa: E402

metric = int(sys.argv[1]) if len(sys.argv) > 1 else 1
layout = sys.argv[2] if len(sys.argv) > 2 else "shuffled"
rng = np.random.default_rng(9)
d = int(sys.argv[3]) if len(sys.argv) > 3 else 128
k = int(sys.argv[4]) if len(sys.argv) > 4 else 10
storage = m.STORE_BF16 if len(sys.argv) > 5 and sys.argv[5] == "bf16" else m.STORE_F32
nq = int(sys.argv[6]) if len(sys.argv) > 6 else 512
nc, dup = {"huge": (6, 4000), "dense": (60, 400)}.get(layout, (300, 80))
centres = rng.standard_normal((nc, d)).astype(np.float32)
xb = (np.repeat(centres, dup, axis=0) + rng.standard_normal((nc * dup, d)).astype(np.float32) * 1e-3).astype(np.float32)
if layout != "ordered":
    xb = xb[rng.permutation(len(xb))]
xq = (centres[rng.integers(0, nc, nq)] + rng.standard_normal((nq, d)).astype(np.float32) * 1e-3).astype(np.float32)
ix = m.IndexFlat(d, metric, storage=storage)
ix.add(xb)
exact = m.IndexFlat(d, metric, storage=storage)
exact.add(xb)
De, Ie = exact.set_search_params(algo=m.ALGO_SCAN).search(xq, k)
ix.set_search_params(algo=m.ALGO_TENSOR, profile=True)
keys = ("fallback_queries", "overflow_queries", "rescued_queries", "range_queries", "last_kprime", "last_list_entries",
        "last_main_ms", "last_total_ms", "last_launches")
print(layout, "metric", metric, "d", d, "k", k, "storage", "bf16" if storage == m.STORE_BF16 else "f32", "nq", nq)
# distances of the exact scan recomputed in float64 on the authoritative rows: what "same ids" may differ by among ties
rows = ix.reconstruct_n().astype(np.float64)
def true_d(I):
    g = rows[np.maximum(I, 0)]
    q = xq.astype(np.float64)[:, None, :]
    return ((g - q) ** 2).sum(-1) if metric == 1 else (g * q).sum(-1)
De64 = true_d(Ie)
# ground truth: float64 brute force over the authoritative rows
allk = (((rows ** 2).sum(1)[None, :] - 2.0 * xq.astype(np.float64) @ rows.T + (xq.astype(np.float64) ** 2).sum(1)[:, None])
        if metric == 1 else -(xq.astype(np.float64) @ rows.T))
kth = np.sort(allk, axis=1)[:, k - 1]                 # true k-th best key per query
def audit(name, D_, I_):
    td = true_d(I_)
    key = td if metric == 1 else -td
    worse = (key - kth[:, None]).max()                 # > 0: a returned row is worse than the true k-th by that much
    dup = max(len(r_) - len(set(r_.tolist())) for r_ in I_)
    print(f"   {name}: max |returned D - float64 D of the returned id| {np.abs(D_ - td).max():.3g} (rel {np.abs(D_ - td).max() / max(1e-30, np.abs(td).max()):.2g});"
          f" worst returned row vs the true k-th key: {worse:+.3g}; duplicate ids in a result row: {dup}", flush=True)
audit("exact scan", De, Ie)
for i in range(5):
    t0 = time.perf_counter()
    D, I = ix.search(xq, k)
    dt = time.perf_counter() - t0
    st = ix.stats()
    print(i, f"{dt * 1e3:.2f} ms", f"vs exact scan: ids equal {float((I == Ie).mean()):.4f}, distances bit-equal {float((D == De).mean()):.4f},",
          f"max |dD| {float(np.abs(D - De).max()):.3g}, worst float64 distance gap at differing ids {float(np.abs(true_d(I) - De64)[I != Ie].max()) if (I != Ie).any() else 0.0:.3g}",
          {k_: st[k_] for k_ in ("fallback_queries", "overflow_queries", "rescued_queries", "range_queries", "last_kprime", "last_total_ms")}, flush=True)
    if i in (0, 2):
        audit("tensor path", D, I)
