"""Times the ingest-side kernels on one B200: K5 (fp32 rows -> rows + bf16 scan copy + norms) through
index.add() with device tensors, and K7 (CLS / masked-mean pooling + optional normalise, fused with the ingest
writes) through index.add_pooled().  Prints one JSON line per case with achieved GB/s of algorithmic bytes."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import rag_faiss_embedding_b200 as b2f  # noqa: E402
from rag_faiss_embedding_b200.encoder import synth_rows  # noqa: E402


def timed(fn, reps=5):
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best


def main():
    peak = 6452.8
    try:
        peak = float(json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")))["hbm_gbs"])
    except Exception:
        pass
    d = 384
    for n in (1_000_000, 5_000_000):
        x = synth_rows(1234, 0, n, d)
        for storage, name in ((b2f.STORE_F32, "fp32"), (b2f.STORE_BF16, "bf16")):
            ix = b2f.IndexFlat(d, 1, storage=storage)
            ix.reserve(n)

            def run():
                ix.reset()
                ix.add(x)

            ms = timed(run)
            byts = n * d * (4 + (4 if storage == b2f.STORE_F32 else 0) + 2) + n * 4
            print(json.dumps({"kernel": "K5 ingest (index.add, device rows)", "rows": n, "d": d, "storage": name, "ms": round(ms, 4),
                              "GBps": round(byts / ms / 1e6, 1), "frac_of_hbm_peak": round(byts / ms / 1e6 / peak, 3),
                              "rows_per_s": round(n / ms * 1e3)}), flush=True)
            del ix
        del x
    B, T = 8192, 128
    hidden = torch.randn(B, T, d, device="cuda")
    mask = (torch.rand(B, T, device="cuda") < 0.7).to(torch.int64)
    mask[:, 0] = 1
    for pool, byts in (("cls", B * d * 4 * 2 + B * d * 2 + B * 4), ("mean", B * T * d * 4 + B * T * 8 + B * d * 4 + B * d * 2 + B * 4)):
        ix = b2f.IndexFlat(d, 1)
        ix.reserve(B)

        def run():
            ix.reset()
            ix.add_pooled(hidden, mask, pool=pool, normalize=True)

        ms = timed(run)
        print(json.dumps({"kernel": f"K7 pool({pool}) + normalise + ingest (index.add_pooled)", "B": B, "T": T, "d": d, "ms": round(ms, 4),
                          "GBps": round(byts / ms / 1e6, 1), "frac_of_hbm_peak": round(byts / ms / 1e6 / peak, 3),
                          "chunks_per_s": round(B / ms * 1e3)}), flush=True)


if __name__ == "__main__":
    main()
