"""BASELINE configs[4] (default --chunks 5000000: its stated size): end-to-end RAG ingest on 1..N B200s -- a random-init MiniLM-L6 encoder
(transformers BertModel: 6 layers, hidden 384, 12 heads, intermediate 1536, vocab 30522; there is no network for
the real weights) embeds synthetic chunks (random token ids, lengths ~U[16,128]); the encoder output stays on the
device and goes through the fused pooling + L2-normalise + add kernel (K7) straight into index storage; then a
query batch is encoded the same way and searched (k = 10).  One process per GPU (torchrun): every rank encodes
and stores its own share of the chunks (row shards), queries are replicated, per-shard results are merged after
one all-gather.  The encoder is torch (out of scope of this repo); K7 / add / search are ours.

    python tools/c5_pipeline.py --chunks 200000 --queries 10000          # miniature
    python -m torch.distributed.run --nproc-per-node 8 --master-addr 127.0.0.1 tools/c5_pipeline.py   # 5M chunks, 10k queries

Prints one JSON line: encode chunks/s, K7+add time and GB/s, search q/s, fallbacks, and a parity check of sampled
queries against a torch fp32 brute force over the rows every shard holds.
"""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def synth_batch(gen, batch, dev):
    lens = torch.randint(16, 129, (batch,), generator=gen, device=dev)
    T = int(lens.max().item())
    ids = torch.randint(1000, 30000, (batch, T), generator=gen, device=dev)
    mask = (torch.arange(T, device=dev)[None, :] < lens[:, None]).to(torch.int64)
    return ids * mask, mask


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--chunks", type=int, default=5_000_000, help="chunks per job (split over the ranks)")
    ap.add_argument("--check", type=int, default=32, help="queries checked against a torch fp32 brute force")
    ap.add_argument("--queries", type=int, default=10_000)
    ap.add_argument("--batch", type=int, default=2048)
    ap.add_argument("--pool", default="mean", choices=["mean", "cls"])
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    import torch.distributed as dist

    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    from transformers import BertConfig, BertModel

    import rag_faiss_embedding_b200 as b2f
    from rag_faiss_embedding_b200.encoder import pool_normalize
    from rag_faiss_embedding_b200.sharded import partition_rows

    torch.manual_seed(0)
    cfg = BertConfig(vocab_size=30522, hidden_size=384, num_hidden_layers=6, num_attention_heads=12,
                     intermediate_size=1536, max_position_embeddings=512)
    model = BertModel(cfg, add_pooling_layer=False).to(dev).eval()
    d = 384
    lo, hi = partition_rows(args.chunks, world)[rank]
    sh = b2f.ShardedIndexFlat(d, b2f.METRIC_INNER_PRODUCT, device=local)
    sh.local.reserve(hi - lo)
    gen = torch.Generator(device=dev)
    gen.manual_seed(1234 + rank)
    ev = lambda: torch.cuda.Event(enable_timing=True)  # noqa: E731
    enc_ms = add_ms = 0.0
    k7_bytes = 0
    done = 0
    with torch.inference_mode(), torch.autocast("cuda", dtype=torch.bfloat16):
        while done < hi - lo:
            b = min(args.batch, hi - lo - done)
            ids, mask = synth_batch(gen, b, dev)
            e0, e1, e2 = ev(), ev(), ev()
            e0.record()
            hidden = model(input_ids=ids, attention_mask=mask).last_hidden_state.float()
            e1.record()
            sh.local.add_pooled(hidden, mask, pool=args.pool, normalize=True)   # K7: no host bounce
            e2.record()
            torch.cuda.synchronize()
            if done > 0:   # skip the first (warm-up) batch in the rates
                enc_ms += e0.elapsed_time(e1)
                add_ms += e1.elapsed_time(e2)
                T = hidden.shape[1]
                k7_bytes += (b * T * d * 4 + b * T * 8 if args.pool == "mean" else b * d * 4) + b * d * 6 + b * 4
            done += b
        sh.segments.append(lo, hi - lo)
        sh.set_total(args.chunks)
        # queries: encoded once (replicated), pooled by the same kernel, searched in one batch
        genq = torch.Generator(device=dev)
        genq.manual_seed(99)
        qs = []
        for q0 in range(0, args.queries, args.batch):
            ids, mask = synth_batch(genq, min(args.batch, args.queries - q0), dev)
            hidden = model(input_ids=ids, attention_mask=mask).last_hidden_state.float()
            qs.append(pool_normalize(hidden, mask, args.pool, True))
        xq = torch.cat(qs)
    for _ in range(2):
        sh.search(xq, 10)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    e0, e1 = ev(), ev()
    e0.record()
    reps = 5
    for _ in range(reps):
        D, I = sh.search(xq, 10)
    e1.record()
    torch.cuda.synchronize()
    search_ms = e0.elapsed_time(e1) / reps
    # ---- parity of sampled queries: torch fp32 brute force over the rows each shard actually holds, merged ----------
    nchk = min(args.check, args.queries)
    recall = None
    if nchk > 0:
        rows = torch.from_numpy(sh.local.reconstruct_n()).to(dev)          # the authoritative rows of this shard
        torch.backends.cuda.matmul.allow_tf32 = False
        ip = xq[:nchk].float() @ rows.T                                      # fp32 matmul
        kk = min(10, rows.shape[0])
        v, idx = torch.topk(ip, kk, dim=1)
        idx = idx + lo
        if world > 1:
            vs = [torch.empty_like(v) for _ in range(world)]
            ids = [torch.empty_like(idx) for _ in range(world)]
            dist.all_gather(vs, v)
            dist.all_gather(ids, idx)
            v, idx = torch.cat(vs, 1), torch.cat(ids, 1)
        top_v, pos = torch.topk(v, 10, dim=1)
        top_i = torch.gather(idx, 1, pos)
        got = I[:nchk]
        hits = sum(len(set(top_i[r].tolist()) & set(got[r].tolist())) for r in range(nchk))
        # ids may differ only among (near-)ties of the fp32 brute force itself: compare the distances too
        recall = {"queries": nchk, "recall_at_10": hits / (10.0 * nchk),
                  "max_abs_ip_diff": float((top_v - D[:nchk]).abs().max())}
        del rows, ip
    t = torch.tensor([enc_ms, add_ms, search_ms], device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    enc_ms, add_ms, search_ms = [float(v) for v in t.tolist()]
    if rank == 0:
        timed_chunks = (hi - lo) - min(args.batch, hi - lo)
        st = sh.local.stats()
        print(json.dumps({
            "config": "BASELINE configs[4]%s: random-init MiniLM-L6 encode + fused pool/normalise/add, then a %d-query search"
                      % ("" if args.chunks >= 5_000_000 else " in miniature", args.queries),
            "n_gpus": world, "chunks_total": args.chunks, "chunks_per_gpu": hi - lo, "pool": args.pool, "queries": args.queries,
            "encode_chunks_per_s_per_gpu": round(timed_chunks / enc_ms * 1e3, 1), "encode_dtype": "bf16 autocast (torch; not this repo's code)",
            "k7_add_ms_total": round(add_ms, 3), "k7_add_GBps": round(k7_bytes / add_ms / 1e6, 1),
            "k7_share_of_ingest": round(add_ms / (enc_ms + add_ms), 5),
            "search_ms": round(search_ms, 4), "search_qps": round(args.queries / search_ms * 1e3, 1),
            "search_algo": {1: "scan", 2: "tensor"}.get(st["last_algo"]), "fallback_queries": st["fallback_queries"],
            "self_check": {"top1_ip_min": round(float(D[:, 0].min()), 4), "labels_in_range": bool(((I >= 0) & (I < args.chunks)).all())},
            "parity_vs_torch_fp32_bruteforce": recall, "rescued_queries": st["rescued_queries"],
            "range_queries": st["range_queries"], "searches": st["searches"], "last_kprime": st["last_kprime"],
        }), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
