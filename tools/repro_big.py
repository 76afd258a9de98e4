import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import oracle as orc
import rag_faiss_embedding_b200 as b2f
d = 384
n = int(sys.argv[1]); nq = int(sys.argv[2]); storage = b2f.STORE_BF16 if sys.argv[3] == "bf16" else b2f.STORE_F32
ix = b2f.IndexFlat(d, 1, storage=storage); ix.reserve(n); ix.add_synthetic(1234, 0, n)
xq = orc.c_synth_rows(5678, 0, nq, d)
ix.set_search_params(algo=b2f.ALGO_TENSOR, profile=True)
for it in range(3):
    try:
        D, I = ix.search(xq, 10)
        s = ix.stats()
        print(f"n={n} nq={nq} {sys.argv[3]} it={it}: ok survivors/q={s['last_list_entries']/nq:.1f} overflow={s['overflow_queries']} fallback={s['fallback_queries']} kernel_ms={s['last_main_ms']:.3f}", flush=True)
    except Exception as e:
        print(f"n={n} nq={nq} {sys.argv[3]} it={it}: FAIL {e}", flush=True)
        break
# self-query sanity: rows of the database must find themselves at distance 0
rows = [0, n // 2, n - 1]
q = np.stack([orc.c_synth_rows(1234, r, 1, d)[0] for r in rows])
if storage == b2f.STORE_BF16:
    import torch
    q = torch.from_numpy(q).to(torch.bfloat16).to(torch.float32).numpy()
try:
    D, I = ix.search(np.concatenate([q, xq[:61]]), 10)
    print("self-query ids", I[:3, 0].tolist(), "want", rows, "d0", D[:3, 0].tolist(), flush=True)
except Exception as e:
    print("self-query FAIL", e, flush=True)
