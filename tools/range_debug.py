"""Diagnostics for the failure paths behind the tensor pass on near-duplicate data: four searches, stats after each.

usage: range_debug.py [metric] [layout]
  layout "shuffled": 300 clusters x 80 near-duplicates in random row order  -> uncertified queries WITH a k-th distance
                     (the range pass serves them from the second search on)
         "ordered" : the same rows, every cluster stored contiguously       -> thresholds never tighten, lists overflow
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
d, k, nq = 128, 10, 512
nc, dup = {"huge": (6, 4000), "dense": (60, 400)}.get(layout, (300, 80))
centres = rng.standard_normal((nc, d)).astype(np.float32)
xb = (np.repeat(centres, dup, axis=0) + rng.standard_normal((nc * dup, d)).astype(np.float32) * 1e-3).astype(np.float32)
if layout != "ordered":
    xb = xb[rng.permutation(len(xb))]
xq = (centres[rng.integers(0, nc, nq)] + rng.standard_normal((nq, d)).astype(np.float32) * 1e-3).astype(np.float32)
ix = m.IndexFlat(d, metric)
ix.add(xb)
exact = m.IndexFlat(d, metric)
exact.add(xb)
De, Ie = exact.set_search_params(algo=m.ALGO_SCAN).search(xq, k)
ix.set_search_params(algo=m.ALGO_TENSOR, profile=True)
keys = ("fallback_queries", "overflow_queries", "rescued_queries", "range_queries", "last_kprime", "last_list_entries",
        "last_main_ms", "last_total_ms", "last_launches")
print(layout, "metric", metric)
for i in range(5):
    t0 = time.perf_counter()
    D, I = ix.search(xq, k)
    dt = time.perf_counter() - t0
    st = ix.stats()
    print(i, f"{dt * 1e3:.2f} ms", f"vs exact scan: ids equal {float((I == Ie).mean()):.4f}, distances bit-equal {float((D == De).mean()):.4f},",
          f"max |dD| {float(np.abs(D - De).max()):.3g}", {k_: st[k_] for k_ in keys}, flush=True)
