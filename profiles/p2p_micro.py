"""Latency of the two exchanges of the sharded CRD step: NVLink peer-memory kernels vs NCCL, payloads of the headline
config (B=46 anchors, D=128).   torchrun --nproc-per-node N profiles/p2p_micro.py"""
import os, sys, json, torch
import torch.distributed as dist
sys.path.insert(0, '.')
import __graft_entry__ as ge
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
dev = torch.device("cuda", int(os.environ["LOCAL_RANK"]))
dist.init_process_group("nccl", device_id=dev)
pkg = ge.load_package()
from crdpn_b200.sharded import PeerExchange
B, D = 46, 128
counts = [B * (r + 1) // world - B * r // world for r in range(world)]
px = PeerExchange(None, rank, world, dev, 64, 128)
v1 = torch.randn(counts[rank], D, device=dev); v2 = torch.randn(counts[rank], D, device=dev)
y = torch.arange(counts[rank], device=dev) + 1000 * rank
packed = torch.randn(2 * B * D + 8, device=dev)
rows = max(counts)
buf = torch.randn(rows, 2 * D + 2, device=dev); gout = torch.empty(world * rows, 2 * D + 2, device=dev)

def timeit(fn, iters=300, graph=False):
    for _ in range(20): fn()
    torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
    run = fn
    if graph:
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, capture_error_mode="thread_local"):
            for _ in range(10): fn()
        run = g.replay; iters //= 10
        run(); torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters): run()
    e1.record(); torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1) / (iters * (10 if graph else 1)) * 1e3], device=dev, dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return round(t.item(), 2)

res = {"world": world}
# correctness of the exchange itself
g1, g2, gy = px.allgather(v1, v2, y, counts)
ref = [torch.empty(2 * B * D + 8, device=dev) for _ in range(world)]
dist.all_gather(ref, packed)
want = ref[0].clone()
for r in range(1, world): want += ref[r]
got = px.allreduce(packed)
res["allreduce_bitwise_rank_order"] = bool(torch.equal(got, want))
a0 = sum(counts[:rank])
res["allgather_own_rows_ok"] = bool(torch.equal(g1[a0:a0 + counts[rank]], v1) and torch.equal(gy[a0:a0 + counts[rank]], y))
for graph in (False, True):
    tag = "graph" if graph else "eager"
    res[f"p2p_allgather_us_{tag}"] = timeit(lambda: px.allgather(v1, v2, y, counts), graph=graph)
    res[f"p2p_allreduce_us_{tag}"] = timeit(lambda: px.allreduce(packed), graph=graph)
    res[f"nccl_allgather_us_{tag}"] = timeit(lambda: dist.all_gather_into_tensor(gout, buf), graph=graph)
    res[f"nccl_allreduce_us_{tag}"] = timeit(lambda: dist.all_reduce(packed), graph=graph)
if rank == 0:
    print(json.dumps(res))
torch.cuda.synchronize(); dist.barrier()
px.close()
dist.destroy_process_group()
