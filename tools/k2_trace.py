"""Where K2's time goes: per-CTA %globaltimer stamps from the diagnostics build of the library (tools/k2_trace.sh),
summarised per phase and per logical unit (segments, tiles, when its MMA stream ended).
usage: k2_trace.py [rows] [d] [nq ...]"""
import ctypes
import os
import sys

HERE = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
os.environ["B200FLAT_LIB"] = os.path.join(HERE, "rag-faiss-embedding_b200", "lib", "libb200flat_trace.so")
sys.path.insert(0, HERE)
import numpy as np  # noqa: E402
import torch  # noqa: E402

import rag_faiss_embedding_b200 as m  # noqa: E402
from rag_faiss_embedding_b200 import _capi  # noqa: E402

rows = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
d = int(sys.argv[2]) if len(sys.argv) > 2 else 384
nqs = [int(a) for a in sys.argv[3:]] or [32, 1024]
SL = 12
lib = _capi.load()
lib.b2f_debug_k2_trace.argtypes = [ctypes.c_void_p, ctypes.c_int]
ix = m.IndexFlat(d, m.METRIC_L2)
ix.add_synthetic(1234, 0, rows)
ix.set_search_params(algo=m.ALGO_TENSOR, profile=True)
names = ["entry", "setup done", "first tile landed", "last MMA issued", "last tile examined", "lists pruned", "all warps done", "TMEM freed"]
flush = torch.empty(512 << 20, dtype=torch.uint8, device="cuda")
buf = np.zeros(296 * SL, dtype=np.uint64)
for nq in nqs:
    xq = torch.randn(nq, d, device="cuda")
    for _ in range(5):
        ix.search(xq, 10)
    flush.zero_()
    torch.cuda.synchronize()
    assert lib.b2f_debug_k2_trace(buf.ctypes.data, buf.size) == 0   # (reading clears the stamps)
    ix.search(xq, 10)
    torch.cuda.synchronize()
    st = ix.stats()
    assert lib.b2f_debug_k2_trace(buf.ctypes.data, buf.size) == 0
    t = buf.reshape(296, SL).astype(np.int64)
    t = t[t[:, 0] > 0]
    t0 = t[:, 0].min()
    rel = np.where(t > 0, (t - t0) / 1e3, np.nan)
    lead = t[:, 3] > 0   # CTAs that issue MMAs (every CTA, or the leader of every pair)
    print(f"nq {nq}: {len(t)} CTAs ({int(lead.sum())} issue MMAs), K2 by CUDA events {st['last_main_ms'] * 1e3:.1f} us, first entry -> last exit {np.nanmax(rel[:, 6:8]):.1f} us")
    for s_, nm in enumerate(names):
        col = rel[:, s_][t[:, s_] > 0]
        if len(col):
            print(f"   {nm:20s} min {col.min():8.1f}  p10 {np.percentile(col, 10):8.1f}  median {np.median(col):8.1f}  p90 {np.percentile(col, 90):8.1f}  max {col.max():8.1f} us  ({len(col)} CTAs)")
    mma = rel[:, 3][lead]
    first = rel[:, 2][lead]
    print(f"   MMA streams: start {np.median(first):.1f}, end median {np.median(mma):.1f}, ends spread over {mma.max() - mma.min():.1f} us; "
          f"last MMA issued -> kernel end {np.nanmax(rel[:, 6:8]) - mma.max():.1f} us")
    # per logical unit: the plan's segments next to the measured stream
    import ctypes as C
    out = (C.c_int32 * 12)()
    lib.b2f_plan_describe(nq, rows, d, 10, 0, out)
    kp, units, T, R = out[0], out[5], (out[10] + (1 if out[4] else 0)) // (2 if out[4] else 1), out[11]
    ntiles = (rows + 255) // 256
    print(f"   plan: k' {kp}, {units} units over {T} query tile units, rounds of {R} database tiles, {ntiles} database tiles")
    rowsu = []
    for r_ in range(len(t)):
        if not lead[r_]:
            continue
        u = int(t[r_, 8]) - 1
        seg = (C.c_int32 * 14)()
        cnt = (C.c_int32 * 2)()
        ns = lib.b2f_plan_unit_work(T, units, R, ntiles, kp, u, seg, None, 0, cnt)
        rowsu.append((rel[r_, 3], u, int(t[r_, 9]), ns, cnt[0], cnt[1] if ns > 1 else 0, seg[0], seg[7] if ns > 1 else -1,
                      rel[r_, 10] if t[r_, 10] > 0 else float("nan"), rel[r_, 2]))
    rowsu.sort()
    print("   MMA-end us | unit | SM | segments | tiles s0 s1 | query tile s0 s1 | s0 MMAs issued at | first tile at")
    for e in rowsu[:6] + [None] + rowsu[-12:]:
        if e is None:
            print("      ...")
        else:
            print(f"   {e[0]:9.1f}  {e[1]:4d} {e[2]:4d}  {e[3]}   {e[4]:5d} {e[5]:5d}   {e[6]:3d} {e[7]:3d}   {e[8]:8.1f}  {e[9]:6.1f}")
    a = np.array([(e[0], e[3], e[4] + e[5], e[2]) for e in rowsu])
    for ns in (1, 2):
        sel = a[:, 1] == ns
        if sel.any():
            print(f"   units with {ns} segment(s): {int(sel.sum())}, MMA end mean {a[sel, 0].mean():.1f} us, tiles mean {a[sel, 2].mean():.1f}")
    die = a[:, 3] >= 74
    print(f"   MMA end by SM id: < 74: {a[~die, 0].mean():.1f} us ({int((~die).sum())}),  >= 74: {a[die, 0].mean():.1f} us ({int(die.sum())});  even SM {a[a[:, 3] % 2 == 0, 0].mean():.1f}, odd SM {a[a[:, 3] % 2 == 1, 0].mean() if (a[:, 3] % 2 == 1).any() else float('nan'):.1f}")
    print(f"   correlation(MMA end, tiles) = {np.corrcoef(a[:, 0], a[:, 2])[0, 1]:.2f}")
