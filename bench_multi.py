"""Multi-GPU leg of bench.py (launched by torchrun, one rank per GPU, NCCL for set-up only).

STRONG scaling of BASELINE.json configs[3] as written: B = 46 anchors, K = 65536 negatives, N = 1M rows x 2 banks,
fp32, banks row-sharded over R = 2 / 4 / 8 GPUs (rank r owns rows [N r / R, N (r+1) / R)), replicated contrast_idx.
Per step and rank: peer-memory all-gather of the local anchors' embeddings -> fused score + loss + backward pass over the
entries of contrast_idx that live in this shard -> ONE kernel that reduces the partials, momentum-updates the owned
positive rows and sums gradients / loss over the ranks (LL words over NVLink).  3 launches, no NCCL call, captured in
a CUDA graph.  value = 2 B (K+1) / max-over-ranks device time: total work is fixed, so "scaling": "strong".

Parity is asserted INSIDE this run (the driver's box runs the 2-GPU pytest cases only when it has two GPUs): loss and
gradients of the sharded step against the unsharded single-GPU step on rank 0 and, for two anchors, against the CPU
oracle (checker only, outside every timed region); updated bank rows bit-identical to the unsharded step's.

"also" keeps the weak-scaling generalisation of round 1 (1M rows + 65536 in-shard negatives per rank) and the PointNet
replicas.
"""
from __future__ import annotations

import ctypes
import os
import sys
import time

from bench import (HEADLINE, SEED, ClockSampler, algorithmic_bytes, make_opt, print_line, scores_per_step, workload_name)


def _global_bank(torch, c):
    """The same [N, 2, D] fp32 bank on every rank (seeded on the host), row-normalised like a warmed bank."""
    g = torch.Generator().manual_seed(SEED + 1)
    bank = torch.rand(c["N"], 2, c["D"], generator=g).mul_(2.0).sub_(1.0)
    bank.div_(bank.norm(dim=2, keepdim=True))
    return bank


def _timed(torch, dist, dev, run, steps):
    torch.cuda.synchronize()
    dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        run()
    e1.record()
    torch.cuda.synchronize()
    dist.barrier()
    ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    return ms.item() / steps


def _capture(torch, step, warm, rank):
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        for _ in range(warm):
            step()
    torch.cuda.current_stream().wait_stream(side)
    torch.cuda.synchronize()
    if os.environ.get("CRDPN_NO_GRAPH") is not None:
        return step, None
    try:  # thread_local: NCCL's watchdog thread polls events while we capture
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph, capture_error_mode="thread_local"):
            step()
        return graph.replay, graph
    except Exception as exc:  # keep measuring, eagerly, and say so
        print(f"[bench_multi] CUDA graph capture failed on rank {rank}: {exc}", file=sys.stderr)
        torch.cuda.synchronize()
        return step, None


def _strong(args, pkg, torch, dist, dev, rank, world, hbm_peak, peak_kind):
    c = dict(HEADLINE)
    B, D, K, N = c["B"], c["D"], c["K"], c["N"]
    K1 = K + 1
    opt = make_opt(c)
    torch.manual_seed(SEED)
    crit = pkg.ShardedCRDLoss(opt, comm="p2p", fixed_local_batch=True).to(dev)
    mem = crit.contrast
    mem._peer_exchange(B, D, dev)   # collective set-up (CUDA IPC)
    lo, hi = mem.row_begin, mem.row_end
    bank = _global_bank(torch, c)
    with torch.no_grad():
        mem.memory_v1.copy_(bank[lo:hi, 0]); mem.memory_v2.copy_(bank[lo:hi, 1])
    g = torch.Generator().manual_seed(SEED)
    f_s = torch.randn(B, c["s_dim"], generator=g)
    f_t = torch.randn(B, c["t_dim"], generator=g)
    y = torch.randperm(N, generator=g)[:B]
    cidx = torch.randint(0, N, (B, K1), generator=g)
    cidx[:, 0] = y
    counts = [B * (r + 1) // world - B * r // world for r in range(world)]
    a0 = sum(counts[:rank])
    sl = slice(a0, a0 + counts[rank])
    for p_ in list(crit.embed_s.parameters()) + list(crit.embed_t.parameters()):  # same heads on every rank
        dist.broadcast(p_.data, src=0)
    f_s_d, f_t_d, y_d, cidx_d = f_s.to(dev), f_t.to(dev), y.to(dev), cidx.to(dev)
    with torch.no_grad():
        v1_all = crit.embed_s(f_s_d).contiguous()
        v2_all = crit.embed_t(f_t_d).contiguous()
    v1, v2, y_loc = v1_all[sl].contiguous(), v2_all[sl].contiguous(), y_d[sl].contiguous()
    mem._ensure_counts(counts[rank], dev)
    # first call: Z over all shards (general path), then the banks back to their initial state for the parity check
    g1, g2, gy = mem._gather(v1, v2, y_loc)
    mem._freeze_z(g1, g2, cidx_d)
    hp = mem._host_params()

    # ---- parity of ONE sharded step (from the initial banks) -------------------------------------------------------
    out = mem.step_resident(v1, v2, y_loc, cidx_d)
    torch.cuda.synchronize()
    red = out["reduced"].clone()
    parity = {}
    ok = torch.ones(1, device=dev)
    rel = lambda a, b: ((a.double() - b.double()).abs().max() / (b.double().abs().max() + 1e-300)).item()
    ref_rows = None
    if rank == 0:
        # unsharded single-GPU step on the full bank with the same Z (the path the 1-GPU tests pin to the oracle)
        torch.manual_seed(SEED)
        one = pkg.CRDLoss(opt).to(dev)
        with torch.no_grad():
            one.contrast.memory_v1.copy_(bank[:, 0]); one.contrast.memory_v2.copy_(bank[:, 1])
        res, r1, r2 = one.contrast._step(v1_all, v2_all, y_d, cidx_d, hp.Z1, hp.Z2)
        torch.cuda.synchronize()
        want_loss = (res[0] + res[1]).item()
        parity["loss_rel_vs_unsharded"] = abs(red[2 * B * D + 5].item() - want_loss) / abs(want_loss)
        parity["grad_v1_rel_vs_unsharded"] = rel(red[:B * D].view(B, D), r1)
        parity["grad_v2_rel_vs_unsharded"] = rel(red[B * D:2 * B * D].view(B, D), r2)
        ref_rows = torch.stack([one.contrast.memory_v1[y_d], one.contrast.memory_v2[y_d]], 1).contiguous()
        # CPU oracle (checker only): two anchors, all K+1 entries, full bank
        from oracle import crd_oracle
        crd_oracle.build()
        sub = [0, B - 1]
        o = crd_oracle.crd_score(bank[:, 0].contiguous().numpy(), bank[:, 1].contiguous().numpy(), v1_all[sub].cpu().numpy(),
                                 v2_all[sub].cpu().numpy(), cidx[sub].numpy(), N, c["T"], hp.Z1, hp.Z2)
        scale = B / len(sub)   # the oracle normalises by its own batch of 2 anchors
        og1 = torch.from_numpy(o["grad_v1"]) / scale
        og2 = torch.from_numpy(o["grad_v2"]) / scale
        parity["grad_v1_rel_vs_oracle_2anchors"] = rel(red[:B * D].view(B, D)[sub].cpu(), og1)
        parity["grad_v2_rel_vs_oracle_2anchors"] = rel(red[B * D:2 * B * D].view(B, D)[sub].cpu(), og2)
        del one
        torch.cuda.empty_cache()
        if max(parity.values()) > 1e-4:
            ok.zero_()
    # updated rows: every owner compares its positives with the unsharded step's rows, bit for bit
    rows = torch.empty(B, 2, D, device=dev)
    if rank == 0:
        rows.copy_(ref_rows)
    dist.broadcast(rows, src=0)
    mine = [(i, int(v)) for i, v in enumerate(y.tolist()) if lo <= v < hi]
    bit_ok = all(torch.equal(mem.memory_v1[v - lo], rows[i, 0]) and torch.equal(mem.memory_v2[v - lo], rows[i, 1]) for i, v in mine)
    bits = torch.tensor([1.0 if bit_ok else 0.0], device=dev)
    dist.all_reduce(bits, op=dist.ReduceOp.MIN)
    dist.all_reduce(ok, op=dist.ReduceOp.MIN)
    parity["updated_rows_bit_identical"] = bool(bits.item() == 1.0)
    # identical bits of the reduced buffer on every rank (rank-ordered sums)
    chk = red[:2 * B * D].view(torch.int32).to(torch.int64).sum().reshape(1)
    cmin, cmax = chk.clone(), chk.clone()
    dist.all_reduce(cmin, op=dist.ReduceOp.MIN); dist.all_reduce(cmax, op=dist.ReduceOp.MAX)
    parity["reduced_identical_on_all_ranks"] = bool(cmin.item() == cmax.item())
    parity_ok = bool(ok.item() == 1.0) and parity["updated_rows_bit_identical"] and parity["reduced_identical_on_all_ranks"]

    # ---- device-resident timing -------------------------------------------------------------------------------------
    def step():
        mem.step_resident(v1, v2, y_loc, cidx_d, out)

    run, graph = _capture(torch, step, max(args.warmup, 3), rank)
    for _ in range(3):
        run()
    lib = pkg._native.lib()
    sampler = ClockSampler(dev.index)
    sampler.start()
    ms_step = _timed(torch, dist, dev, run, args.steps)
    clocks = sampler.stop()
    # per-launch duration of the dominant kernel: event-bracketed eager launches right after the timed region
    tot, n = ctypes.c_double(), ctypes.c_uint64()
    lib.crdpn_timing_enable(1)
    lib.crdpn_timing_read(0, ctypes.byref(tot), ctypes.byref(n))
    for _ in range(20):
        step()
    torch.cuda.synchronize()
    lib.crdpn_timing_read(0, ctypes.byref(tot), ctypes.byref(n))
    lib.crdpn_timing_enable(0)
    kms = torch.tensor([tot.value / max(n.value, 1)], dtype=torch.float64, device=dev)
    dist.all_reduce(kms, op=dist.ReduceOp.MAX)
    if graph is not None:
        graph.reset()
    del run, graph

    # ---- end to end through the public API, pinned HOST inputs on every rank -----------------------------------------
    # (a) negatives drawn on each GPU inside its own shard, K / R per rank (no index traffic at all);
    # (b) the replicated [B, K+1] int64 list comes from the host on every rank, every step
    del crit, mem
    torch.cuda.empty_cache()
    K_loc = K // world
    opt_l = make_opt(dict(c, K=K_loc))
    torch.manual_seed(SEED)
    crit_l = pkg.ShardedCRDLoss(opt_l, local_negatives=True, comm="p2p", fixed_local_batch=True).to(dev)
    torch.manual_seed(SEED)
    crit_r = pkg.ShardedCRDLoss(opt, comm="p2p", fixed_local_batch=True).to(dev)
    for cr in (crit_l, crit_r):
        for p_ in list(cr.embed_s.parameters()) + list(cr.embed_t.parameters()):
            dist.broadcast(p_.data, src=0)
    host = [t.pin_memory() for t in (f_s[sl].contiguous(), f_t[sl].contiguous(), y[sl].contiguous(), cidx)]

    def e2e_run(cr, n_in, graphed):
        def e2e_step():
            dev_in = [t.to(dev, non_blocking=True) for t in host[:n_in]]
            dev_in[0].requires_grad_()
            loss = cr(dev_in[0], dev_in[1], dev_in[2], dev_in[3] if n_in == 4 else None)
            cr.zero_grad(set_to_none=True)  # the reference's order: forward, zero_grad, backward
            loss.backward()
            return loss.item()

        for _ in range(3):
            e2e_step()
        gs = None
        if graphed:   # forward + backward of the sharded step captured once; replays stay in lock-step on every rank
            cr.contrast.device_sampler_offset()

            def fn(*dev_in):
                loss = cr(dev_in[0], dev_in[1], dev_in[2], dev_in[3] if n_in == 4 else None)
                loss.backward()
                return loss

            gs = pkg.GraphedStep(fn, host[:n_in], dev, grad_inputs=(0,), zero_grad=lambda: cr.zero_grad(set_to_none=True))

        def loop(n):
            if gs is None:
                for _ in range(n):
                    e2e_step()
                return
            gs.stage(*host[:n_in])
            for i in range(n):
                gs.run()
                if i + 1 < n:
                    gs.stage(*host[:n_in])
                if gs.pending() > 1:
                    gs.collect()
            while gs.pending():
                gs.collect()

        loop(3)
        torch.cuda.synchronize()
        dist.barrier()
        t0 = time.perf_counter()
        esteps = max(args.steps, 20) if graphed else max(args.steps // 4, 5)
        loop(esteps)
        torch.cuda.synchronize()
        dist.barrier()
        ms_ = torch.tensor([(time.perf_counter() - t0) * 1e3 / esteps], dtype=torch.float64, device=dev)
        dist.all_reduce(ms_, op=dist.ReduceOp.MAX)
        return ms_.item(), sum(t.numel() * t.element_size() for t in host[:n_in])

    def guarded(cr, n_in, graphed):
        try:
            return e2e_run(cr, n_in, graphed)
        except Exception as exc:   # keep the line: report the failure instead of losing the run
            print(f"[bench_multi] e2e leg (n_in={n_in}, graphed={graphed}) failed on rank {rank}: {exc}", file=sys.stderr)
            return float("nan"), 0

    e2e_strict_ms, h2d = e2e_run(crit_l, 3, False)
    e2e_strict_ms_h, h2d_h = e2e_run(crit_r, 4, False)
    e2e_ms, _ = guarded(crit_l, 3, True)
    e2e_ms_h, _ = guarded(crit_r, 4, True)
    if e2e_ms != e2e_ms:
        e2e_ms = e2e_strict_ms
    if e2e_ms_h != e2e_ms_h:
        e2e_ms_h = e2e_strict_ms_h
    scores_l = 2 * B * (K_loc * world + 1)
    del crit_l, crit_r
    torch.cuda.empty_cache()

    total_scores = scores_per_step(c)
    per_rank_bytes = (2 * B * K1 * D * 4) / world + B * K1 * 8 + 16 * B * D   # rows of this shard + the whole index list
    achieved = per_rank_bytes / (kms.item() * 1e-3) / 1e9
    line = {
        "metric": "crd_negatives_scored_per_sec", "value": total_scores / (ms_step * 1e-3), "unit": "scores/s",
        "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_step,
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": workload_name(c, world), "B": B, "D": D, "K": K, "N": N, "banks": 2,
                   "parallelism": f"bank-shard{world}", "rows_per_rank": hi - lo,
                   "bank_layout": "interleaved [N/R,2,D] fp32 per rank", "contrast_idx": "replicated [B,K+1] int64, fixed",
                   "l2": f"per-rank banks {2 * (hi - lo) * D * 4 / 1e6:.0f} MB, random rows; no flush",
                   "step": "peer-memory all-gather(anchors) + score/loss/backward over the shard + ONE kernel: reduction, "
                           "momentum update, sum over ranks (LL words over NVLink); 3 launches in a CUDA graph"},
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": hbm_peak, "unit": "GB/s", "frac": achieved / hbm_peak,
                     "traffic": None, "peak_kind": peak_kind, "kernel": "crd_score_kernel (per rank, slowest rank)",
                     "kernel_ms": kms.item(), "algorithmic_bytes": per_rank_bytes,
                     "note": "per-rank algorithmic bytes = 1/R of the row gathers + the whole replicated index list"},
        "parity": dict(parity, ok=parity_ok,
                       what="one sharded step from the initial banks vs the unsharded single-GPU step (all anchors) and the "
                            "CPU oracle (2 anchors), tolerance 1e-4; updated rows and the reduced buffer bit-compared"),
        "e2e": {"value": scores_l / (e2e_ms * 1e-3), "unit": "scores/s", "h2d_bytes_per_step": h2d * world,
                "d2h_bytes_per_step": 4 * world, "ms_per_step": e2e_ms,
                "api": f"GraphedStep over ShardedCRDLoss(f_s_loc, f_t_loc, idx_loc) + backward(): forward and backward captured once in a "
                       f"CUDA graph on every rank; per step: pinned host features + indices in (next batch staged on a copy stream), K/R = "
                       f"{K_loc} in-shard negatives drawn on each GPU (fresh on every replay), the loss read back one step late",
                "sync_each_step": {"value": scores_l / (e2e_strict_ms * 1e-3), "ms_per_step": e2e_strict_ms,
                                   "note": "the reference loop verbatim through the per-step Python path: copy, forward, backward, loss.item()"},
                "with_host_contrast_idx": {"value": total_scores / (e2e_ms_h * 1e-3), "unit": "scores/s",
                                           "h2d_bytes_per_step": h2d_h * world, "ms_per_step": e2e_ms_h,
                                           "sync_each_step": {"value": total_scores / (e2e_strict_ms_h * 1e-3), "ms_per_step": e2e_strict_ms_h},
                                           "note": "the replicated [B,K+1] int64 list copied to EVERY rank each step (staged, GraphedStep)"}},
        "gpu_launches": 3 * args.steps,
        "comm": "NVLink peer-memory kernels (all-gather; all-reduce fused into the reduction kernel), no NCCL call in the step",
        "collectives_per_step": 0, "exchange_kernels_per_step": 1,
        "cuda_graph": os.environ.get("CRDPN_NO_GRAPH") is None,
        "clocks": clocks,
    }
    return line, parity_ok


def _weak(args, pkg, torch, dist, dev, rank, world):
    """Round 1's weak-scaling generalisation: every rank owns 1M rows and scores B = 46 anchors against 65536 negatives
    drawn inside its own shard (global N = R x 1M, global K = R x 65536)."""
    c = dict(HEADLINE)
    K_loc, N_loc, B, D = c["K"], c["N"], c["B"], c["D"]
    cg = dict(c, N=N_loc * world)
    torch.manual_seed(SEED + rank)
    crit = pkg.ShardedCRDLoss(make_opt(cg), local_negatives=True, comm="p2p", fixed_local_batch=True).to(dev)
    mem = crit.contrast
    mem._peer_exchange(B, D, dev)
    lo, hi = mem.row_begin, mem.row_end
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
    for p_ in list(crit.embed_s.parameters()) + list(crit.embed_t.parameters()):
        dist.broadcast(p_.data, src=0)
    f_s_d, f_t_d, y_d, cidx_d = f_s[sl].to(dev), f_t[sl].to(dev), y[sl].to(dev), cidx.to(dev)
    with torch.no_grad():
        v1 = crit.embed_s(f_s_d).contiguous()
        v2 = crit.embed_t(f_t_d).contiguous()
    mem._ensure_counts(counts[rank], dev)
    g1, g2, gy = mem._gather(v1, v2, y_d)
    mem._freeze_z(g1, g2, cidx_d)
    out = mem.step_resident(v1, v2, y_d, cidx_d)

    def step():
        mem.step_resident(v1, v2, y_d, cidx_d, out)

    run, graph = _capture(torch, step, 3, rank)
    for _ in range(3):
        run()
    ms_step = _timed(torch, dist, dev, run, args.steps)
    if graph is not None:
        graph.reset()
    del run, graph, crit, mem
    torch.cuda.empty_cache()
    total = 2 * B * (K_loc * world + 1)
    return {"scaling": "weak", "workload": f"1M rows + {K_loc} in-shard negatives per rank (global N = {world}M, K = {K_loc * world})",
            "value": total / (ms_step * 1e-3), "unit": "scores/s", "ms_per_step": ms_step}


def run_multi(args, pkg, torch, dist, dev, rank, world, hbm_peak, peak_kind):
    import faulthandler
    faulthandler.dump_traceback_later(240, exit=True, file=sys.stderr)  # a stuck exchange must not eat the box time
    os.environ.setdefault("CRDPN_P2P_TIMEOUT_S", "30")  # bench only: fail fast if the ranks' call sequences diverge
    line, parity_ok = _strong(args, pkg, torch, dist, dev, rank, world, hbm_peak, peak_kind)
    also = {}
    try:
        also["weak_scaling"] = _weak(args, pkg, torch, dist, dev, rank, world)
    except Exception as exc:
        also["weak_scaling"] = {"error": str(exc)}
    try:  # BASELINE metric's second half at N > 1: PointNet points/s (replicas, no collective)
        from bench_pointnet import bench_pointnet_replicas
        also["pointnet"] = bench_pointnet_replicas(pkg, torch, dist, dev, rank, world, max(args.steps // 2, 20), args.warmup)
    except Exception as exc:
        also["pointnet"] = {"error": str(exc)}
    line["also"] = also
    if rank == 0:
        print_line(line)
    faulthandler.cancel_dump_traceback_later()
    import threading
    sys.stdout.flush()
    killer = threading.Timer(20.0, lambda: os._exit(0 if parity_ok else 3))
    killer.daemon = True
    killer.start()
    torch.cuda.synchronize()
    dist.barrier()
    dist.destroy_process_group()
    killer.cancel()
    if not parity_ok:
        raise SystemExit("bench_multi: the sharded step does NOT match the unsharded step (see \"parity\" in the JSON line)")
