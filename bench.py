#!/usr/bin/env python
"""bench.py -- queries/sec of the flat-search hot path on B200 (BASELINE.json metric).

Workload at every N: BASELINE config 2 -- synthetic 1M x 384 fp32 rows, flat L2, 1024 queries, k=10.
A "step" is one index.search() over the whole query batch.

  value      device-resident: queries / results are CUDA tensors, timed with CUDA events
  e2e        the same search through the C-ABI with HOST (pinned) buffers: H2D of the queries and D2H
             of (D, I) are inside the timed region
  roofline   the dominant kernel (K2 tcgen05 tensor scan; K1 streaming scan when --algo scan) timed
             per launch with CUDA events on its own stream inside the library
  cpu_baseline  the CPU restatement of IndexFlat (oracle/) on the box's host cores, bounded sample

N > 1 (torchrun): the 1M rows are row-sharded over the ranks (contiguous ranges) and the query batch
grows with N (1024 x N queries, replicated on every rank), so per-GPU work is constant ("weak"): every
rank searches its shard for the whole batch, per-rank top-k lists are exchanged with one NCCL all-gather
and merged by the CUDA merge kernel.  value = all queries answered / max-over-ranks time.  --impl reference times the CPU restatement instead (faiss-cpu itself cannot be installed
here: no wheel, no network -- see DESIGN.md).
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: (rows, d, nq, k, metric, normalize, storage)
    "c2": dict(n=1_000_000, d=384, nq=1024, k=10, metric=1, normalize=False, storage="fp32",
               label="synthetic 1Mx384 fp32 flat L2, 1024 queries, k=10 (BASELINE configs[1])"),
    "c2_nq1": dict(n=1_000_000, d=384, nq=1, k=10, metric=1, normalize=False, storage="fp32",
                   label="synthetic 1Mx384 fp32 flat L2, 1 query, k=10 (small-batch series)"),
    "c2_nq4": dict(n=1_000_000, d=384, nq=4, k=10, metric=1, normalize=False, storage="fp32",
                   label="synthetic 1Mx384 fp32 flat L2, 4 queries, k=10 (small-batch series)"),
    "c2_nq8": dict(n=1_000_000, d=384, nq=8, k=10, metric=1, normalize=False, storage="fp32",
                   label="synthetic 1Mx384 fp32 flat L2, 8 queries, k=10 (small-batch series)"),
    "c2_nq128": dict(n=1_000_000, d=384, nq=128, k=10, metric=1, normalize=False, storage="fp32",
                     label="synthetic 1Mx384 fp32 flat L2, 128 queries, k=10 (small-batch series)"),
    "c2_nq4096": dict(n=1_000_000, d=384, nq=4096, k=10, metric=1, normalize=False, storage="fp32",
                      label="synthetic 1Mx384 fp32 flat L2, 4096 queries, k=10 (large-batch series)"),
    "c2_nq32": dict(n=1_000_000, d=384, nq=32, k=10, metric=1, normalize=False, storage="fp32",
                    label="synthetic 1Mx384 fp32 flat L2, 32 queries, k=10 (small-batch series)"),
    "c3_nq1": dict(n=10_000_000, d=768, nq=1, k=100, metric=0, normalize=True, storage="fp32",
                   label="synthetic 10Mx768 flat inner-product (normalized), batch 1, k=100 (BASELINE configs[2])"),
    "c3_nq32": dict(n=10_000_000, d=768, nq=32, k=100, metric=0, normalize=True, storage="fp32",
                    label="synthetic 10Mx768 flat inner-product (normalized), batch 32, k=100 (BASELINE configs[2])"),
    "c3_nq4096": dict(n=10_000_000, d=768, nq=4096, k=100, metric=0, normalize=True, storage="fp32",
                      label="synthetic 10Mx768 flat inner-product (normalized), batch 4096, k=100 (BASELINE configs[2])"),
    "c4shard": dict(n=12_500_000, d=384, nq=4096, k=10, metric=1, normalize=False, storage="bf16",
                    label="synthetic 12.5Mx384 bf16 flat L2 per GPU (config 4 shard), 4096 queries, k=10"),
    "c4shard_nq1": dict(n=12_500_000, d=384, nq=1, k=10, metric=1, normalize=False, storage="bf16",
                        label="synthetic 12.5Mx384 bf16 flat L2 per GPU (config 4 shard), 1 query, k=10"),
    "c4shard_nq32": dict(n=12_500_000, d=384, nq=32, k=10, metric=1, normalize=False, storage="bf16",
                         label="synthetic 12.5Mx384 bf16 flat L2 per GPU (config 4 shard), 32 queries, k=10"),
}
SEED_DB, SEED_Q = 1234, 5678


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            j = json.load(open(p))
            return dict(hbm=float(j["hbm_gbs"]), tf=float(j["bf16_tflops"]), tf_sus=float(j.get("bf16_tflops_sustained", j["bf16_tflops"])),
                        src="measured")
        except Exception:
            pass
    return dict(hbm=6650.0, tf=1590.0, tf_sus=1400.0, src="fallback")


class ClockSampler:
    """Samples SM clocks / throttle reasons during the timed regions: NVML every ~2 ms (nvidia-smi every
    100 ms as the fallback when the NVML binding is missing)."""
    FIELDS = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")
    NVML_BITS = {"sw_power_cap": 0x4, "hw_slowdown": 0x8, "sw_thermal_slowdown": 0x20, "hw_thermal_slowdown": 0x40}

    def __init__(self, gpu_index: int):
        self.gpu = gpu_index
        self.sm, self.mx, self.reasons = [], [], set()
        self._stop = threading.Event()
        self._t = None
        self.source = "nvidia-smi"

    def _sample_nvml(self, nv, h, mx, get_reasons):
        try:
            self.sm.append(float(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)))
            self.mx.append(mx)
            bits = int(get_reasons(h))
            for name, bit in self.NVML_BITS.items():
                if bits & bit:
                    self.reasons.add(name)
        except Exception:
            pass

    def _run_nvml(self, nv, h):
        mx = float(nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM))
        get_reasons = getattr(nv, "nvmlDeviceGetCurrentClocksEventReasons", None) or nv.nvmlDeviceGetCurrentClocksThrottleReasons
        while True:   # at least one sample even when the whole run is shorter than the thread's start-up
            self._sample_nvml(nv, h, mx, get_reasons)
            if self._stop.wait(0.002):
                break
        self._sample_nvml(nv, h, mx, get_reasons)

    def _run_smi(self):
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        while not self._stop.is_set():
            try:
                out = subprocess.run(["nvidia-smi", f"--id={self.gpu}", f"--query-gpu={self.FIELDS}",
                                      "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5).stdout
                parts = [p.strip() for p in out.strip().split(",")]
                if len(parts) >= 8:
                    self.sm.append(float(parts[0]))
                    self.mx.append(float(parts[1]))
                    for name, val in zip(names, parts[4:8]):
                        if val.lower().startswith("active"):
                            self.reasons.add(name)
            except Exception:
                pass
            self._stop.wait(0.1)

    def start(self):
        try:
            import pynvml as nv

            nv.nvmlInit()
            # NVML enumerates physical devices: map through CUDA_VISIBLE_DEVICES when it lists plain indices
            vis = os.environ.get("CUDA_VISIBLE_DEVICES", "")
            idx = self.gpu
            if vis and all(t.strip().isdigit() for t in vis.split(",")):
                idx = int(vis.split(",")[self.gpu])
            h = nv.nvmlDeviceGetHandleByIndex(idx)
            self.source = "nvml"
            self._t = threading.Thread(target=self._run_nvml, args=(nv, h), daemon=True)
        except Exception:
            self._t = threading.Thread(target=self._run_smi, daemon=True)
        self._t.start()

    def stop(self):
        self._stop.set()
        if self._t:
            self._t.join(timeout=6)
        return {"sm_mhz": statistics.median(self.sm) if self.sm else None, "sm_min_mhz": min(self.sm) if self.sm else None,
                "sm_max_mhz": max(self.mx) if self.mx else None, "reasons": sorted(self.reasons), "samples": len(self.sm),
                "source": self.source}


def cpu_baseline(wl, budget_s=20.0, threads=0):
    """Times the CPU restatement of IndexFlat on a bounded sample of the workload."""
    import oracle as orc

    t0 = time.time()
    n = wl["n"]
    n_cpu = min(n, 1_000_000)
    xb = orc.c_synth_rows(SEED_DB, 0, n_cpu, wl["d"], wl["normalize"])
    cores = orc.c_max_threads() if threads <= 0 else threads
    nq_total = wl["nq"]
    if nq_total < 20:
        nq_s = nq_total
        xq = orc.c_synth_rows(SEED_Q, 0, nq_s, wl["d"], wl["normalize"])
        t = time.time()
        orc.c_search(xb, xq, wl["k"], wl["metric"], algo=0, nthreads=threads)
        dt = time.time() - t
        kind_note = "C restatement, seq path (nq<20: one thread per query, as faiss)"
        cores = min(cores, nq_s)
    else:
        nq_s = min(nq_total, 128)
        xq_all = orc.c_synth_rows(SEED_Q, 0, nq_total, wl["d"], wl["normalize"])
        orc.torch_search_blas(xb[:20000], xq_all[:32], wl["k"], wl["metric"])  # thread-pool / MKL warm-up
        t = time.time()
        orc.torch_search_blas(xb, xq_all[:nq_s], wl["k"], wl["metric"])
        dt = time.time() - t
        # grow the sample while it stays within the budget
        while dt < budget_s / 4 and nq_s < nq_total:
            nq_s = min(nq_total, nq_s * 4)
            t = time.time()
            orc.torch_search_blas(xb, xq_all[:nq_s], wl["k"], wl["metric"])
            dt = time.time() - t
        kind_note = "BLAS (torch CPU/MKL sgemm) restatement of faiss's blas path (sgemm blocks + expanded form + clamp)"
    qps = nq_s / dt * (n_cpu / n)  # rows scale linearly; only != 1 when the host cannot hold n rows
    sample = f"{nq_s} of {nq_total} queries x {n_cpu} of {n} rows, {kind_note}; setup {time.time() - t0 - dt:.1f}s"
    if n_cpu != n:
        sample += " (q/s linearly extrapolated in rows)"
    return {"value": round(qps, 2), "unit": "queries/sec", "cores": int(cores), "kind": "port", "sample": sample}


def run_reference(args, wl, rank, world):
    """--impl reference: the CPU restatement of the reference's IndexFlat path, all host threads."""
    if rank != 0:
        return
    import oracle as orc

    n_cpu = min(wl["n"], 1_000_000)
    xb = orc.c_synth_rows(SEED_DB, 0, n_cpu, wl["d"], wl["normalize"])
    nq_s = wl["nq"] if wl["nq"] < 20 else min(wl["nq"], 1024)   # one step = up to 1024 queries (~1 s of sgemm on 16 cores)
    xq = orc.c_synth_rows(SEED_Q, 0, nq_s, wl["d"], wl["normalize"])

    def step():
        if nq_s < 20:
            orc.c_search(xb, xq, wl["k"], wl["metric"], algo=0)
        else:
            orc.torch_search_blas(xb, xq, wl["k"], wl["metric"])

    for _ in range(max(args.warmup, 1)):
        step()
    t = time.time()
    for _ in range(args.steps):
        step()
    dt = (time.time() - t) / args.steps
    qps = nq_s / dt * (n_cpu / wl["n"])
    cores = orc.c_max_threads()
    line = {
        "impl": "reference", "metric": "queries/sec @k=10 (flat L2, 384-d)", "value": round(qps, 2), "unit": "queries/sec",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(dt * 1e3, 3),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": wl["label"], "rows": wl["n"], "d": wl["d"], "nq": wl["nq"], "k": wl["k"]},
        "cpu_baseline": {"value": round(qps, 2), "unit": "queries/sec", "cores": int(cores), "kind": "port",
                         "sample": f"{nq_s} of {wl['nq']} queries x {n_cpu} rows per step; CPU restatement of IndexFlat "
                                   f"(faiss-cpu unavailable), MKL sgemm blas path" if nq_s >= 20 else
                                   f"{nq_s} queries x {n_cpu} rows per step; C restatement, seq path"},
        "e2e": {"value": round(qps, 2), "unit": "queries/sec", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="c2", choices=sorted(WORKLOADS))
    ap.add_argument("--algo", default="auto", choices=["auto", "scan", "tensor"])
    ap.add_argument("--slack", type=int, default=0, help="tensor path: extra coarse candidates per query (0 = library default)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cpu-budget", type=float, default=20.0)
    args = ap.parse_args()
    wl = WORKLOADS[args.workload]
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))

    if args.impl == "reference":
        run_reference(args, wl, rank, world)
        return

    import torch
    import torch.distributed as dist

    import rag_faiss_embedding_b200 as b2f
    from rag_faiss_embedding_b200 import _capi
    from rag_faiss_embedding_b200.encoder import synth_rows
    from rag_faiss_embedding_b200.sharded import ShardedIndexFlat, partition_rows

    if args.warmup < 3:
        args.warmup = 3
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    peaks = load_peaks()
    algo = {"auto": b2f.ALGO_AUTO, "scan": b2f.ALGO_SCAN, "tensor": b2f.ALGO_TENSOR}[args.algo]
    storage = b2f.STORE_BF16 if wl["storage"] == "bf16" else b2f.STORE_F32

    n, d, nq, k = wl["n"], wl["d"], wl["nq"], wl["k"]
    shard_per_gpu = args.workload.startswith("c4shard")
    if shard_per_gpu:   # per-GPU shard of config 4: every rank holds n rows, batch fixed
        lo, hi = rank * n, (rank + 1) * n
        n_global = n * world
    else:               # the same database row-sharded; the batch grows with N so per-GPU work is constant
        lo, hi = partition_rows(n, world)[rank]
        n_global = n
        nq = nq * world
    ix = b2f.IndexFlat(d, wl["metric"], storage=storage, device=local_rank)
    ix.reserve(hi - lo)
    ix.add_synthetic(SEED_DB, lo, hi - lo, wl["normalize"])   # generated on the device, bit-identical to the oracle
    ix.set_search_params(algo=algo, id_offset=lo, profile=True, slack=args.slack)
    sh = ShardedIndexFlat(d, wl["metric"], local_index=ix) if world > 1 else None
    if sh is not None:
        sh.segments.append(lo, hi - lo)
        sh.set_total(n_global)

    # queries from the library's own counter-based generator (bit-identical to the oracle's): the oracle package is
    # touched only by the checker legs below (parity spot check, cpu_baseline) and by --impl reference
    xq_host = synth_rows(SEED_Q, 0, nq, d, wl["normalize"], device=local_rank).cpu().numpy()
    xq_pin = torch.from_numpy(xq_host).pin_memory()
    D_pin = torch.empty((nq, k), dtype=torch.float32).pin_memory()
    I_pin = torch.empty((nq, k), dtype=torch.int64).pin_memory()
    xq_dev = xq_pin.to(dev)

    def step_device():
        if sh is not None:
            return sh.search(xq_dev, k)
        return ix.search(xq_dev, k)

    def step_host():
        if sh is not None:
            sh.search_host(xq_pin, k, D_pin, I_pin)   # each rank uploads its slice of the replicated batch
        else:
            ix.search_into(xq_pin.numpy(), k, D_pin.numpy(), I_pin.numpy())

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        barrier()
        ms = e0.elapsed_time(e1) / steps
        if world > 1:
            t = torch.tensor([ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms

    # ---- warm-up, then the device-resident timed region -------------------------------------------------
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()   # samples through warm-up, the timed region and the e2e region (GPU busy throughout)
    for _ in range(args.warmup):
        step_device()
    launches0 = ix.stats()["launches"]
    prof0 = ix.stats()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        step_device()
    e1.record()
    barrier()
    ms_step = e0.elapsed_time(e1) / args.steps
    if world > 1:
        t = torch.tensor([ms_step], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_step = float(t.item())
    prof1 = ix.stats()
    # per-launch time of the dominant kernel: CUDA events recorded by the library around every launch in
    # the timed region (running sums, read once so that the host loop stays tight)
    n_launch = max(prof1["prof_main_launches"] - prof0["prof_main_launches"], 1)
    n_search = max(prof1["prof_searches"] - prof0["prof_searches"], 1)
    launches_per_search = n_launch / n_search   # > 1 when a big batch x big k' is processed in query chunks
    main_ms = [(prof1["prof_main_ms_sum"] - prof0["prof_main_ms_sum"]) / n_launch]
    total_ms = [(prof1["prof_total_ms_sum"] - prof0["prof_total_ms_sum"]) / max(prof1["prof_searches"] - prof0["prof_searches"], 1)]
    st = ix.stats()
    launches = st["launches"] - launches0 + (args.steps if world > 1 else 0)  # + the merge kernel per step
    # ---- end-to-end through host buffers ----------------------------------------------------------------
    for _ in range(3):
        step_host()
    e2e_ms = timed(step_host, args.steps)
    clocks = sampler.stop() if rank == 0 else None

    # ---- parity spot check on the timed configuration (not timed): first 4 queries vs the C oracle -------
    parity = None
    if rank == 0 and world == 1 and n <= 1_000_000 and not args.no_cpu_baseline:
        import oracle as orc  # the checker

        xb_host = orc.c_synth_rows(SEED_DB, 0, n, d, wl["normalize"])
        nchk = min(4, nq)
        D_ref, I_ref = orc.c_search(xb_host, xq_host[:nchk], k, wl["metric"], algo=1)
        Dg, Ig = ix.search(xq_host, k)
        parity = orc.recall_and_errors(Dg[:nchk], Ig[:nchk], D_ref, I_ref, wl["metric"])
        del xb_host

    if rank == 0:
        used_algo = st["last_algo"]
        kernel_ms = statistics.mean(main_ms)
        rows_local = hi - lo
        elem = 2 if (used_algo == b2f.ALGO_TENSOR or storage == b2f.STORE_BF16) else 4
        dpad = (d + 63) // 64 * 64 if elem == 2 else d
        if used_algo == b2f.ALGO_TENSOR and nq > 128:
            flops = 2.0 * nq * rows_local * d / launches_per_search   # per launch
            ach = flops / (kernel_ms * 1e-3) / 1e12
            roof = {"bound": "tensor", "achieved": round(ach, 2), "peak": peaks["tf"], "unit": "TFLOP/s",
                    "frac": round(ach / peaks["tf"], 4), "traffic": None}
        else:
            nlaunch = max(st["last_main_launches"], 1)
            passes = (nq + 7) // 8 if used_algo == b2f.ALGO_SCAN else 1   # the scan walks the queries in groups of <= 8
            bytes_ = (rows_local * dpad * elem + min(nq, 8 if used_algo == b2f.ALGO_SCAN else nq) * d * 4) * passes
            ach = bytes_ / (kernel_ms * 1e-3) / 1e9
            roof = {"bound": "hbm", "achieved": round(ach, 1), "peak": peaks["hbm"], "unit": "GB/s",
                    "frac": round(ach / peaks["hbm"], 4), "traffic": None, "launches_per_step": nlaunch}
        try:
            tr = json.load(open(os.path.join(ROOT, "profiles", "ncu_traffic.json")))
            key = f"{args.workload}|{'tensor' if used_algo == b2f.ALGO_TENSOR else 'scan'}"
            if key in tr and world == 1:
                roof["traffic"] = tr[key]["bytes"]
                roof["traffic_source"] = tr[key]["capture"]
        except Exception:
            pass
        roof["kernel"] = "tensor_scan_kernel (K2 tcgen05)" if used_algo == b2f.ALGO_TENSOR else "scan_kernel (K1)"
        roof["kernel_ms"] = round(kernel_ms, 4)
        roof["launches_per_search"] = round(launches_per_search, 2)
        roof["peak_source"] = peaks["src"] + (" burst" if roof["bound"] == "tensor" else "")
        roof["pipeline_ms"] = round(statistics.mean(total_ms), 4)
        line = {
            "metric": "queries/sec @k=%d (flat %s, %d-d)" % (k, "L2" if wl["metric"] == 1 else "IP", d),
            "value": round(nq / (ms_step * 1e-3), 1), "unit": "queries/sec",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(ms_step, 4),
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "bf16 tensor-core candidates + f32 exact re-rank" if used_algo == b2f.ALGO_TENSOR else "f32",
            "data": "synthetic",
            "config": {"workload": wl["label"], "rows_total": n_global, "rows_per_gpu": rows_local, "d": d, "nq": nq, "k": k,
                       "storage": wl["storage"], "algo": {1: "scan", 2: "tensor"}.get(used_algo, "?"),
                       "kprime": st["last_kprime"], "fallback_queries": st["fallback_queries"], "overflow_queries": st["overflow_queries"], "filter_survivors_per_query": round(st["last_list_entries"] / max(nq, 1), 1),
                       "l2_policy": "inputs larger than L2 (bf16 scan copy %.0f MB per GPU, L2 126 MB)" % (rows_local * dpad * 2 / 1e6),
                       "sharding": ("rows in contiguous ranges over %d GPUs, batch %d = %d x N replicated; one packed NCCL all-gather (12 nq k bytes per rank) + CUDA merge per step"
                                    % (world, nq, wl["nq"])) if world > 1 else "single GPU"},
            "roofline": roof,
            "e2e": {"value": round(nq / (e2e_ms * 1e-3), 1), "unit": "queries/sec", "ms_per_step": round(e2e_ms, 4),
                    "h2d_bytes_per_step": nq * d * 4, "d2h_bytes_per_step": nq * k * 12},
            "gpu_launches": int(launches),
            "clocks": clocks,
        }
        if parity is not None:
            line["parity_check"] = parity
        if not args.no_cpu_baseline and world == 1:
            line["cpu_baseline"] = cpu_baseline(wl, args.cpu_budget)
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
