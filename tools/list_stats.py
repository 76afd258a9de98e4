"""GPU diagnostic: LIST-mode filter statistics (survivors per query, overflows) over database / batch sizes."""
import sys, os, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import oracle as orc
import rag_faiss_embedding_b200 as b2f

d = 384
for storage in (b2f.STORE_F32, b2f.STORE_BF16):
    for n in (1_000_000, 4_000_000, 12_500_000):
        if storage == b2f.STORE_F32 and n > 4_000_000:
            continue
        ix = b2f.IndexFlat(d, 1, storage=storage)
        ix.reserve(n)
        ix.add_synthetic(1234, 0, n)
        for nq in (128, 1024, 4096):
            xq = orc.c_synth_rows(5678, 0, nq, d)
            ix.set_search_params(algo=b2f.ALGO_TENSOR, profile=True)
            ix.search(xq, 10)
            s0 = ix.stats()
            t = time.time(); ix.search(xq, 10); dt = time.time() - t
            s = ix.stats()
            print(f"storage={'bf16' if storage else 'f32'} n={n} nq={nq}: survivors/q={s['last_list_entries']/nq:.1f} "
                  f"overflow+={s['overflow_queries']-s0['overflow_queries']} fallback+={s['fallback_queries']-s0['fallback_queries']} "
                  f"kernel_ms={s['last_main_ms']:.3f} total_ms={s['last_total_ms']:.3f} wall_ms={dt*1e3:.2f}", flush=True)
        del ix
