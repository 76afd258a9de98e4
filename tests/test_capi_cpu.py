"""CPU tier: the C-ABI library loads and exports exactly what include/b200flat.h declares; host-side
logic (partitioning, label mapping, the FAISSVectorStore mirror) behaves like the reference's."""
import ctypes
import os
import pickle
import re

import numpy as np
import pytest

import oracle as orc
from tests.helpers import OracleIndex

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "b200flat.h")).read()
    return sorted(set(re.findall(r"B2F_API\s+[\w\s\*]+?\b(b2f_\w+)\s*\(", text)))


def test_header_symbols_exported_and_bound():
    from rag_faiss_embedding_b200 import _capi

    lib = _capi.load()
    declared = _declared_symbols()
    assert len(declared) >= 20
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in b200flat.h but not exported"
    assert sorted(_capi.PROTOTYPES) == declared
    assert lib.b2f_version() == 1
    assert ctypes.sizeof(_capi.SearchParams) == 32 and ctypes.sizeof(_capi.Stats) == 128


def test_no_cpu_fallback_fails_loudly():
    import rag_faiss_embedding_b200 as m

    if m.device_count() > 0:
        pytest.skip("a B200 is present; the failure path is for CPU-only boxes")
    with pytest.raises(m.B200FlatError) as e:
        m.IndexFlatL2(384)
    assert e.value.code == -2 and "no CPU path" in str(e.value)
    with pytest.raises(RuntimeError):
        m.read_index(os.path.join(ROOT, "tests", "golden", "faiss_index.bin"))


def test_faiss_shim_surface():
    import importlib
    import sys

    shim = os.path.join(ROOT, "rag-faiss-embedding_b200", "shim")
    sys.path.insert(0, shim)
    try:
        sys.modules.pop("faiss", None)
        faiss = importlib.import_module("faiss")
        for name in ("IndexFlatL2", "IndexFlatIP", "IndexFlat", "read_index", "write_index", "METRIC_L2",
                     "METRIC_INNER_PRODUCT"):
            assert hasattr(faiss, name)
        assert faiss.METRIC_L2 == 1 and faiss.METRIC_INNER_PRODUCT == 0
    finally:
        sys.path.remove(shim)
        sys.modules.pop("faiss", None)


def test_partition_and_segments():
    from rag_faiss_embedding_b200.sharded import SegmentMap, partition_rows

    assert partition_rows(10, 4) == [(0, 3), (3, 6), (6, 9), (9, 10)]
    assert partition_rows(2, 4) == [(0, 1), (1, 2), (2, 2), (2, 2)]
    assert partition_rows(0, 2) == [(0, 0), (0, 0)]
    for n in (1, 7, 100, 12345):
        for w in (1, 2, 4, 8):
            parts = partition_rows(n, w)
            assert parts[0][0] == 0 and parts[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(parts, parts[1:]))
    seg = SegmentMap()
    seg.append(5, 3)      # local 0..2 -> 5..7
    seg.append(8, 2)      # contiguous: local 3..4 -> 8..9 (same segment)
    seg.append(100, 4)    # local 5..8 -> 100..103
    assert seg.single_offset() is None and len(seg.local_starts) == 2
    got = seg.to_global_numpy(np.array([[0, 2, 4, 5, 8, -1]]))
    assert got.tolist() == [[5, 7, 9, 100, 103, -1]]
    import torch

    got_t = seg.to_global_torch(torch.tensor([[0, 2, 4, 5, 8, -1]]))
    assert got_t.tolist() == [[5, 7, 9, 100, 103, -1]]


def test_store_mirror_host_logic(monkeypatch, tmp_path, golden):
    """FAISSVectorStore mirror: coercions, nq forced to 1, -1 filtering, swallow-all-errors search,
    mapping pickle next to the index, sequential ids when the mapping is missing (faiss_store.py)."""
    from rag_faiss_embedding_b200 import store as st

    written = {}

    def fake_write(index, path):
        orc.np_write_index(path, index.x, index.metric_type)
        written["path"] = path

    def fake_read(path, device=None):
        x, metric = orc.np_read_index(path)
        ix = OracleIndex(x.shape[1], metric)
        ix.add(x)
        return ix

    monkeypatch.setattr(st, "IndexFlatL2", OracleIndex)
    monkeypatch.setattr(st, "IndexFlatIP", lambda d, **kw: OracleIndex(d, orc.METRIC_INNER_PRODUCT))
    monkeypatch.setattr(st, "write_index", fake_write)
    monkeypatch.setattr(st, "read_index", fake_read)

    path = str(tmp_path / "data" / "faiss_index.bin")
    s = st.FAISSVectorStore(dimension=8, index_path=path)
    x = orc.np_synth_rows(3, 0, 6, 8)
    s.add_vectors(x[:5].tolist(), [50, 40, 30, 20, 10])   # list input is coerced to fp32
    s.add_vectors(x[5], [60])                              # 1-D input becomes one row
    assert s.index.ntotal == 6 and s.doc_ids == [50, 40, 30, 20, 10, 60]
    d, ids = s.search(x[2], k=3)
    assert ids[0] == 30 and len(ids) == 3 and d[0] == 0
    d, ids = s.search(x[2].tolist(), k=10)                 # k > ntotal: -1 rows dropped
    assert len(ids) == 6
    d, ids = s.search(np.zeros(5, np.float32), k=3)        # wrong dimension: swallowed, empty result
    assert len(ids) == 0 and d.size == 0
    dm, idm = s.search_many(x[:2], k=2)
    assert [r[0] for r in idm] == [50, 40] and dm.shape == (2, 2)
    s.save_index()
    assert os.path.exists(path) and pickle.load(open(path + ".mapping", "rb")) == s.doc_ids
    s2 = st.FAISSVectorStore(dimension=8, index_path=path)  # auto-load in the constructor
    assert s2.index.ntotal == 6 and s2.doc_ids == s.doc_ids
    os.remove(path + ".mapping")
    s2.load_index()
    assert s2.doc_ids == list(range(6))
    s2.reset()
    assert s2.index.ntotal == 0 and s2.doc_ids == []
    # the reference's own artefacts load through the same surface
    s3 = st.FAISSVectorStore(dimension=384, index_path=golden["index_path"])
    assert s3.doc_ids == golden["mapping"] and s3.index.ntotal == 23
    xb, _ = orc.np_read_index(golden["index_path"])
    case = next(c for c in golden["cases"] if c["metric"] == 1 and c["k"] == 5 and c["queries"] == "self")
    for row in (0, 7, 22):
        d, ids = s3.search(xb[row], k=5)
        assert ids == [golden["mapping"][i] for i in case["ids"][row]]


def test_tensor_path_planner():
    """Host logic of the tensor path (no GPU needed): k', query chunking for feasibility and balance, the
    LIST / HEAP choice, CTA pairs, list geometry."""
    from rag_faiss_embedding_b200 import _capi

    lib = _capi.load()

    def plan(nq, n, d, k, slack=0):
        out = (ctypes.c_int32 * 12)()
        assert lib.b2f_plan_describe(nq, n, d, k, slack, out) == 0
        return dict(zip(["kp", "chunk", "passes", "list", "pair", "units", "ns_min", "nlists", "j", "cap", "tiles", "round"], list(out)))

    c2 = plan(1024, 1_000_000, 384, 10)               # BASELINE config 2
    assert c2["kp"] == 32 and c2["passes"] == 1 and c2["list"] == 1 and c2["pair"] == 1
    assert c2["units"] == 72 and c2["tiles"] == 8 and c2["ns_min"] == 18 and c2["j"] == 1   # 18 whole pairs per pair tile (2 of 74 idle)
    big = plan(4096, 1_000_000, 384, 10)              # 16 pair tiles over 74 pairs: ONE pass, every pair gets 16/74 of it
    assert big["passes"] == 1 and big["chunk"] == 4096 and big["units"] == 74 and big["pair"] == 1 and big["j"] <= 8
    n8 = plan(8192, 125_000, 384, 10)                 # the weak-scaling shape at N = 8: 32 pair tiles over 74 pairs
    assert n8["passes"] == 1 and n8["units"] == 74 and n8["tiles"] == 64 and n8["list"] == 1 and n8["j"] <= 16
    c3 = plan(4096, 10_000_000, 768, 100)             # k' = 192 caps a pass (j <= 16); d = 768 rules out the Q-resident pair kernel
    assert c3["kp"] == 192 and c3["passes"] >= 2 and c3["pair"] == 0 and c3["j"] <= 16 and c3["list"] == 1
    assert c3["units"] == 144 and c3["ns_min"] == 9   # 16 tiles per pass: 9 whole CTAs each beat 148 balanced (A/B r02l)
    one = plan(1, 1_000_000, 384, 10)                 # one tile: every SM streams its own split
    assert one["units"] == 148 and one["passes"] == 1 and one["nlists"] == 296 and one["j"] == 1
    assert plan(64, 1_000_000, 384, 200)["kp"] == 0   # k' > 256: the exact scan serves it
    assert plan(64, 1_000_000, 384, 100, slack=28)["kp"] == 128
    tiny = plan(64, 100, 64, 10)                      # a single database tile: HEAP selection
    assert tiny["kp"] == 32 and tiny["list"] == 0
    small = plan(64, 300, 64, 10)                     # two database tiles: one LIST split per query tile
    assert small["list"] == 1 and small["ns_min"] == 1
    huge = plan(100_000, 1_000_000, 384, 10)          # far more than one wave of query tiles: LIST-mode passes, not HEAP
    assert huge["list"] == 1 and huge["pair"] == 1 and huge["passes"] * huge["chunk"] >= 100_000 and huge["chunk"] <= 74 * 256
    for nq in (2, 127, 129, 300, 1000, 5000, 20000):
        for n in (5_000, 125_000, 12_500_000):
            p = plan(nq, n, 384, 10)
            assert p["kp"] == 32 and p["chunk"] * p["passes"] >= nq
            if p["list"]:   # one wave, shared thresholds need every unit resident
                assert p["units"] <= (74 if p["pair"] else 148) and p["j"] <= 16 and p["cap"] >= 128
                assert p["nlists"] * p["j"] >= 1


def test_balanced_work_split_is_an_exact_partition():
    """The tensor pass's work split, computed by the functions the kernel itself runs (host build of k2::unit_segments /
    k2::SegIter): for every query tile unit the segments' database tiles are every tile exactly once; every unit gets
    the same amount of work (+-2 tiles) whatever the tile / unit counts are; list slots are dense per tile and all
    segments of a tile agree on its voucher slots."""
    from rag_faiss_embedding_b200 import _capi

    lib = _capi.load()

    def unit_work(T, U, R, ntiles, u, cap=60000, kp=32):
        seg, cnt, tiles = (ctypes.c_int32 * 14)(), (ctypes.c_int32 * 2)(), (ctypes.c_int64 * (2 * cap))()
        n = lib.b2f_plan_unit_work(T, U, R, ntiles, kp, u, seg, tiles, cap, cnt)
        assert n in (1, 2), (T, U, u, n)
        return [dict(qtile=seg[7 * s], slot=seg[7 * s + 1], nv=seg[7 * s + 2], p0=seg[7 * s + 3], p1=seg[7 * s + 4],
                     j=seg[7 * s + 5], g=seg[7 * s + 6], tiles=list(tiles[s * cap:s * cap + cnt[s]])) for s in range(n)]

    cases = [(4, 74, 38, 3907), (32, 74, 6, 489), (16, 74, 10, 48829), (1, 148, 296, 3907), (1, 74, 148, 3907), (74, 74, 2, 1000),
             (73, 74, 4, 977), (37, 74, 4, 500), (5, 148, 60, 79), (64, 148, 6, 489), (3, 7, 6, 100), (7, 7, 2, 30),
             (13, 148, 24, 40000), (8, 74, 20, 3907), (2, 3, 4, 9), (1, 1, 2, 5)]
    for T, U, R, ntiles in cases:
        per_tile = {t: [] for t in range(T)}
        slots = {t: {} for t in range(T)}
        totals = []
        for u in range(U):
            segs = unit_work(T, U, R, ntiles, u)
            totals.append(sum(len(s["tiles"]) for s in segs))
            for i, s in enumerate(segs):
                assert s["tiles"] == sorted(s["tiles"])              # swept front to back
                per_tile[s["qtile"]] += s["tiles"]
                assert s["slot"] not in slots[s["qtile"]]
                slots[s["qtile"]][s["slot"]] = s
                if i == 1:   # the piece a unit runs second is the shorter one, sits in the neighbouring tile, never vouches
                    assert s["slot"] >= s["nv"] and abs(s["qtile"] - segs[0]["qtile"]) == 1
                    assert s["p1"] - s["p0"] <= segs[0]["p1"] - segs[0]["p0"]
                    assert segs[0]["slot"] < segs[0]["nv"]           # ... and the longer one, run first, does
        for t in range(T):
            assert sorted(per_tile[t]) == list(range(ntiles)), (T, U, R, ntiles, t)
            ns = len(slots[t])
            assert sorted(slots[t]) == list(range(ns))
            nv = slots[t][0]["nv"]
            assert 1 <= nv <= ns and all(s["nv"] == nv for s in slots[t].values())
            # the tile's voucher lists (2 column halves per piece) together vouch for >= k' rows; j <= 16 register slots;
            # with equal j any g of them do, with unequal pieces all are consulted
            vj = [slots[t][sl]["j"] for sl in range(nv)]
            g = slots[t][0]["g"]
            assert all(j >= 1 for j in vj) and all(s["g"] == g for s in slots[t].values())
            if (T, U) not in ((73, 74), (64, 148)):   # (shapes the planner rejects: a tile with less than one whole unit vouching)
                assert all(j <= 16 for j in vj), (T, U, t, vj)
            assert 2 * sum(vj) >= 32 and 1 <= g <= 2 * nv
            if len(set(vj)) == 1:
                assert g * vj[0] >= 32
            else:
                assert g == 2 * nv
        ideal = T * ntiles / U
        assert max(totals) - min(totals) <= 4 and max(totals) <= ideal + 3, (T, U, R, ntiles, min(totals), max(totals), ideal)
    # plans the library makes are consistent with the split: the lists allocated per query cover every tile's segments
    out = (ctypes.c_int32 * 12)()
    for nq, n in ((1024, 1_000_000), (4096, 1_000_000), (8192, 125_000), (4096, 12_500_000), (300, 50_000), (128, 1_000_000)):
        assert lib.b2f_plan_describe(nq, n, 384, 10, 0, out) == 0
        kp, units, nlists, tiles, pair, rnd = out[0], out[5], out[7], out[10], out[4], out[11]
        T = tiles // 2 if pair else tiles
        ntiles = (n + 255) // 256
        most = 0
        for u in range(units):
            for s in unit_work(T, units, rnd, ntiles, u, cap=1, kp=kp):
                most = max(most, s["slot"] + 1)
        assert most * 2 == nlists, (nq, n, most, nlists)


def test_bench_reference_arm_contract():
    """`bench.py --impl reference` (the CPU restatement timed on the host cores) prints ONE JSON line with the keys
    the driver reads; runs without a GPU."""
    import json
    import subprocess
    import sys

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "1",
                          "--workload", "c2_nq1"], capture_output=True, text=True, timeout=300, cwd=root)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [ln for ln in out.stdout.splitlines() if ln.startswith("{")]
    assert len(lines) == 1
    j = json.loads(lines[0])
    assert j["impl"] == "reference" and j["unit"] == "queries/sec" and j["higher_is_better"] is True and j["value"] > 0
    assert j["steps"] == 1 and j["n_gpus"] == 1 and j["vs_baseline"] is None and j["data"] == "synthetic"
    assert j["cpu_baseline"]["kind"] == "port" and j["cpu_baseline"]["cores"] >= 1 and j["cpu_baseline"]["value"] == j["value"]
    assert j["e2e"] == {"value": j["value"], "unit": "queries/sec", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in j["config"]
    # both arms print the SAME workload-defining config (one function), so the driver's same_config comparison holds
    import importlib.util

    spec = importlib.util.spec_from_file_location("bench_mod", os.path.join(root, "bench.py"))
    bench = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(bench)
    assert j["config"] == bench.workload_config(bench.WORKLOADS["c2_nq1"], 1)
    c8 = bench.workload_config(bench.WORKLOADS["c2"], 8)
    assert c8["nq"] == 8192 and c8["rows_per_gpu"] == 125000 and c8["l2_policy"].startswith("L2 flushed")   # 96 MB < L2
    assert bench.workload_config(bench.WORKLOADS["c2"], 1)["l2_policy"].startswith("inputs larger than L2")
    c4 = bench.workload_config(bench.WORKLOADS["c4shard"], 8)
    assert c4["rows_total"] == 100_000_000 and c4["rows_per_gpu"] == 12_500_000 and c4["nq"] == 4096
    assert j["cpu_baseline"]["one_thread"]["cores"] == 1

