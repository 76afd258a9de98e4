"""2-GPU diagnostic: where does the time go in a sharded search step (search vs all_gather vs merge)?"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.distributed as dist
import oracle as orc
import rag_faiss_embedding_b200 as b2f

rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"]); lr = int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr); dev = torch.device("cuda", lr)
dist.init_process_group("nccl", device_id=dev)
d, n, nq, k = 384, 1_000_000, 1024, 10
ix = b2f.ShardedIndexFlat(d, 1, device=lr); ix.add_synthetic(1234, n)
xq = torch.from_numpy(orc.c_synth_rows(5678, 0, nq, d)).to(dev)

def t(fn, iters=20):
    for _ in range(3): fn()
    torch.cuda.synchronize(); dist.barrier(); t0 = time.time()
    for _ in range(iters): fn()
    torch.cuda.synchronize(); return (time.time() - t0) / iters * 1e3

D, I = ix.search_local(xq, k)
Dg = torch.empty((world * nq, k), dtype=D.dtype, device=dev); Ig = torch.empty((world * nq, k), dtype=I.dtype, device=dev)
res = {
    "local_search_ms": t(lambda: ix.search_local(xq, k)),
    "allgather_D_ms": t(lambda: dist.all_gather_into_tensor(Dg, D)),
    "allgather_I_ms": t(lambda: dist.all_gather_into_tensor(Ig, I)),
    "merge_ms": t(lambda: b2f.merge_topk(1, Dg.view(world, nq, k), Ig.view(world, nq, k))),
    "full_search_ms": t(lambda: ix.search(xq, k)),
}
big = torch.empty(64 << 20, dtype=torch.uint8, device=dev); bigg = torch.empty(world * (64 << 20), dtype=torch.uint8, device=dev)
ms = t(lambda: dist.all_gather_into_tensor(bigg, big), 5)
res["allgather_64MB_ms"] = ms; res["allgather_64MB_GBps_per_rank"] = (world - 1) * 64 / 1024 / (ms / 1e3)
if rank == 0: print(res, flush=True)
dist.destroy_process_group()
