"""GPU diagnostic (not a test): runs the tensor path on a ladder of shapes with certification off and
prints recall / error statistics against the float64 oracle, so one gpurun call localises a K2 bug
(descriptor / swizzle / pipeline-phase / selection)."""
import sys
import os
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import oracle as orc  # noqa: E402
import rag_faiss_embedding_b200 as b2f  # noqa: E402

SHAPES = [
    # n, d, nq, k
    (256, 64, 128, 10), (256, 64, 1, 10), (256, 128, 128, 10), (512, 64, 128, 10), (1024, 384, 128, 10),
    (5000, 384, 128, 10), (5000, 384, 300, 10), (100000, 384, 1024, 10), (100000, 768, 256, 10),
]


def main():
    for metric in (1, 0):
        for (n, d, nq, k) in SHAPES:
            xb = orc.c_synth_rows(1234, 0, n, d)
            xq = orc.c_synth_rows(5678, 0, nq, d)
            ix = b2f.IndexFlat(d, metric)
            ix.add(xb)
            ix.set_search_params(algo=b2f.ALGO_TENSOR, certify=False, profile=True)
            t = time.time()
            try:
                D, I = ix.search(xq, k)
            except Exception as e:  # noqa: BLE001
                print(f"metric={metric} n={n} d={d} nq={nq}: EXCEPTION {e}", flush=True)
                continue
            dt = time.time() - t
            D_ref, I_ref = orc.np_search_f64(xb, xq, k, metric)
            r = orc.recall_and_errors(D, I, D_ref, I_ref, metric)
            st = ix.stats()
            per_q = [(len(set(I[q]) & set(I_ref[q]))) for q in range(nq)]
            bad_q = [q for q in range(nq) if per_q[q] < min(k, n)]
            print(f"metric={metric} n={n} d={d} nq={nq} k={k}: recall={r['recall']:.4f} idmis={r['id_mismatch']} "
                  f"relerr={r['max_rel_err']:.2e} main_ms={st['last_main_ms']:.3f} total_ms={st['last_total_ms']:.3f} "
                  f"wall={dt*1e3:.1f}ms bad_queries={bad_q[:8]}{'...' if len(bad_q) > 8 else ''}", flush=True)
            if bad_q:
                q = bad_q[0]
                print("   q", q, "got", I[q].tolist(), "want", I_ref[q].tolist())
                print("   D got", np.round(D[q], 3).tolist(), "want", np.round(D_ref[q], 3).tolist())
            # with certification on, results must be exact whatever the coarse pass did
            ix.set_search_params(certify=True)
            D, I = ix.search(xq, k)
            r2 = orc.recall_and_errors(D, I, D_ref, I_ref, metric)
            print(f"   certified: recall={r2['recall']:.4f} idmis={r2['id_mismatch']} fallback_queries={ix.stats()['fallback_queries']} overflow={ix.stats()['overflow_queries']}",
                  flush=True)


if __name__ == "__main__":
    main()
