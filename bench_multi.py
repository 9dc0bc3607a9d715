"""Multi-GPU leg of bench.py (launched by torchrun, one rank per GPU, NCCL).

Weak scaling of the CRD step (BASELINE.json configs[3] generalised): every rank owns a 1M-row shard of both banks
(global N = R x 1M) and scores the B=46 anchors of the global batch against K_loc = 65536 negatives drawn inside its
own shard (global K = R x 65536).  Per step: one packed all-gather of the anchors' embeddings, one fused
score+loss+backward pass per rank over its shard, one packed all-reduce of the partials, owner-only momentum update.
value = scores of ALL ranks / max-over-ranks device time.
"""
from __future__ import annotations

import ctypes
import json
import os
import time

from bench import (HEADLINE, SEED, ClockSampler, algorithmic_bytes, make_opt, print_line, scores_per_step, workload_name)


def run_multi(args, pkg, torch, dist, dev, rank, world, hbm_peak, peak_kind):
    import faulthandler
    import sys
    faulthandler.dump_traceback_later(120, exit=True, file=sys.stderr)  # a stuck collective must not eat the box time
    c = dict(HEADLINE)
    K_loc, N_loc, B, D = c["K"], c["N"], c["B"], c["D"]
    cg = dict(c, N=N_loc * world)                       # global bank
    opt = make_opt(cg)
    torch.manual_seed(SEED + rank)
    # exchanges: "p2p" = single kernels over NVLink peer memory (default), "nccl" = torch.distributed collectives
    comm = os.environ.get("CRDPN_COMM", "p2p")
    crit = pkg.ShardedCRDLoss(opt, local_negatives=True, comm="p2p" if comm == "p2p" else "dist").to(dev)
    mem = crit.contrast
    if comm == "p2p":  # collective set-up; if CUDA IPC is unavailable on ANY rank, every rank says so and uses NCCL
        ok = torch.ones(1, device=dev)
        try:
            mem._peer_exchange(B, D, dev)
        except Exception as exc:
            print(f"[bench_multi] rank {rank}: peer-memory exchange unavailable ({exc}); using NCCL", file=sys.stderr)
            ok.zero_()
        dist.all_reduce(ok, op=dist.ReduceOp.MIN)
        if ok.item() == 0:
            comm, mem.comm = "nccl", "dist"
    lo, hi = mem.row_begin, mem.row_end
    # global batch, identical on every rank (seeded); each rank embeds its slice of the anchors
    g = torch.Generator().manual_seed(SEED)
    f_s = torch.randn(B, c["s_dim"], generator=g)
    f_t = torch.randn(B, c["t_dim"], generator=g)
    y = torch.randperm(cg["N"], generator=g)[:B]
    counts = [B * (r + 1) // world - B * r // world for r in range(world)]
    a0 = sum(counts[:rank])
    sl = slice(a0, a0 + counts[rank])
    gl = torch.Generator().manual_seed(SEED * 1000 + rank)
    cidx = torch.randint(lo, hi, (B, K_loc + 1), generator=gl)
    cidx[:, 0] = y
    for p_ in list(crit.embed_s.parameters()) + list(crit.embed_t.parameters()):  # same heads on every rank
        dist.broadcast(p_.data, src=0)
    f_s_d, f_t_d, y_d, cidx_d = f_s[sl].to(dev), f_t[sl].to(dev), y[sl].to(dev), cidx.to(dev)
    with torch.no_grad():
        v1 = crit.embed_s(f_s_d).contiguous()
        v2 = crit.embed_t(f_t_d).contiguous()

    def step():
        g1, g2, gy = mem._gather(v1, v2, y_d)                      # exchange 1
        mem._freeze_z(g1, g2, cidx_d)
        hp = mem._host_params()
        res, d1, d2 = mem._step(g1, g2, gy, cidx_d, hp.Z1, hp.Z2)   # local shard
        return mem._reduce_partials(res, d1, d2)                   # exchange 2

    # warm up on a side stream, then capture one whole step (2 collectives + 2 kernels + glue) in a CUDA graph:
    # at 0.45 ms of GPU work per step the host-side launch cost of ~25 small ops is otherwise exposed
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        for _ in range(max(args.warmup, 3)):
            step()
    torch.cuda.current_stream().wait_stream(side)
    torch.cuda.synchronize()
    use_graph = os.environ.get("CRDPN_NO_GRAPH") is None
    run = step
    if use_graph:
        try:  # thread_local: NCCL's watchdog thread polls events while we capture
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph, capture_error_mode="thread_local"):
                step()
            run = graph.replay
        except Exception as exc:  # keep measuring, eagerly, and say so
            print(f"[bench_multi] CUDA graph capture failed on rank {rank}: {exc}", file=sys.stderr)
            use_graph = False
            torch.cuda.synchronize()
    for _ in range(3):
        run()
    lib = pkg._native.lib()
    tot, n = ctypes.c_double(), ctypes.c_uint64()
    l0 = pkg._native.launch_count()
    sampler = ClockSampler(dev.index)
    sampler.start()
    torch.cuda.synchronize()
    dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        run()
    e1.record()
    torch.cuda.synchronize()
    dist.barrier()
    clocks = sampler.stop()
    ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms_step = ms.item() / args.steps
    # per-launch duration of the dominant kernel: event-bracketed eager launches of the same step right after the
    # timed region (event records cannot live inside the captured graph)
    lib.crdpn_timing_enable(1)
    lib.crdpn_timing_read(0, ctypes.byref(tot), ctypes.byref(n))
    for _ in range(20):
        step()
    torch.cuda.synchronize()
    lib.crdpn_timing_read(0, ctypes.byref(tot), ctypes.byref(n))
    lib.crdpn_timing_enable(0)
    kms = torch.tensor([tot.value / max(n.value, 1)], dtype=torch.float64, device=dev)
    dist.all_reduce(kms, op=dist.ReduceOp.MAX)
    launches = (pkg._native.launch_count() - l0) if not use_graph else (4 if comm == "p2p" else 2) * args.steps  # graph replays

    # end to end through the public API with pinned HOST inputs on every rank: (a) the negatives are drawn on the GPU
    # inside each rank's shard (contrast_idx=None, the module's default), (b) every rank's index list comes from the host
    host = [t.pin_memory() for t in (f_s[sl].contiguous(), f_t[sl].contiguous(), y[sl].contiguous(), cidx)]

    def e2e_run(n_in):
        def e2e_step():
            dev_in = [t.to(dev, non_blocking=True) for t in host[:n_in]]
            dev_in[0].requires_grad_()
            loss = crit(dev_in[0], dev_in[1], dev_in[2], dev_in[3] if n_in == 4 else None)
            crit.zero_grad(set_to_none=True)  # the reference's order: forward, zero_grad, backward
            loss.backward()
            return loss.item()

        for _ in range(3):
            e2e_step()
        torch.cuda.synchronize()
        dist.barrier()
        t0 = time.perf_counter()
        esteps = max(args.steps // 4, 5)
        for _ in range(esteps):
            e2e_step()
        torch.cuda.synchronize()
        dist.barrier()
        ms_ = torch.tensor([(time.perf_counter() - t0) * 1e3 / esteps], dtype=torch.float64, device=dev)
        dist.all_reduce(ms_, op=dist.ReduceOp.MAX)
        return ms_, sum(t.numel() * t.element_size() for t in host[:n_in])

    e2e_ms, h2d = e2e_run(3)
    e2e_ms_h, h2d_h = e2e_run(4)

    total_scores = 2 * B * (K_loc * world + 1)
    per_rank_bytes = algorithmic_bytes(c)
    achieved = per_rank_bytes / (kms.item() * 1e-3) / 1e9
    also = {}
    try:  # BASELINE metric's second half at N > 1: PointNet points/s (replicas, no collective)
        from bench_pointnet import bench_pointnet_replicas
        also["pointnet"] = bench_pointnet_replicas(pkg, torch, dist, dev, rank, world, max(args.steps // 2, 20), args.warmup)
    except Exception as exc:
        also["pointnet"] = {"error": str(exc)}
    if rank == 0:
        line = {
            "metric": "crd_negatives_scored_per_sec", "value": total_scores / (ms_step * 1e-3), "unit": "scores/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_step,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": workload_name(c, world), "B": B, "D": D, "K_per_rank": K_loc, "K_global": K_loc * world,
                       "N_per_rank": N_loc, "N_global": N_loc * world, "banks": 2, "parallelism": f"bank-shard{world}",
                       "bank_layout": "interleaved [N_loc,2,D] fp32 per rank",
                       "l2": "inputs larger than L2 (1.02 GB of banks per rank, random rows); no flush",
                       "step": "all-gather(anchors) + crdpn_crd_step on the local shard + packed all-reduce(partials)"},
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": hbm_peak, "unit": "GB/s", "frac": achieved / hbm_peak,
                         "traffic": None, "peak_kind": peak_kind, "kernel": "crd_score_kernel (per rank, slowest rank)",
                         "kernel_ms": kms.item(), "algorithmic_bytes": per_rank_bytes},
            "e2e": {"value": total_scores / (e2e_ms.item() * 1e-3), "unit": "scores/s", "h2d_bytes_per_step": h2d * world,
                    "d2h_bytes_per_step": 4 * world, "ms_per_step": e2e_ms.item(),
                    "api": "ShardedCRDLoss(f_s_loc, f_t_loc, idx_loc).backward(): pinned host features + indices in, in-shard "
                           "negatives drawn on each GPU, loss.item() out",
                    "with_host_contrast_idx": {"value": total_scores / (e2e_ms_h.item() * 1e-3), "unit": "scores/s",
                                               "h2d_bytes_per_step": h2d_h * world, "ms_per_step": e2e_ms_h.item()}},
            "gpu_launches": launches, "comm": ("NVLink peer-memory kernels (all-gather + one-shot all-reduce), no NCCL call in the step"
                                               if comm == "p2p" else "nccl"),
            "collectives_per_step": 0 if comm == "p2p" else 2, "exchange_kernels_per_step": 2 if comm == "p2p" else 0,
            "cuda_graph": use_graph,
            "clocks": clocks,
            "also": also,
        }
        print_line(line)
    faulthandler.cancel_dump_traceback_later()
    # teardown: drop the captured graph before the communicator; never let a stuck teardown eat box time
    import threading
    sys.stdout.flush()
    killer = threading.Timer(20.0, lambda: os._exit(0))
    killer.daemon = True
    killer.start()
    if use_graph:
        graph.reset()
        del graph, run
    torch.cuda.synchronize()
    dist.barrier()
    dist.destroy_process_group()
    killer.cancel()
