"""PointNet half of the benchmark (BASELINE.json configs[1]): eval-mode encoder, B=160, P=2500, 3->64->128->1024."""
from __future__ import annotations

import ctypes
import os
import time

FLOP_PER_POINT = 2 * (3 * 64 + 64 * 128 + 128 * 1024)  # 278 912


def bench_pointnet(pkg, torch, dev, args, tf_peak, peak_kind, B=160, P=2500, F=1024):
    from oracle import pointnet_oracle as po
    st = po.random_state(F, seed=46)
    enc = pkg.ShapeEncoderPC(F)
    enc.load_state_dict(st)
    enc = enc.to(dev).eval()
    x_host = po.random_clouds(B, P, seed=46).pin_memory()
    x = x_host.to(dev)
    lib = pkg._native.lib()
    steps, warmup = max(args.steps // 2, 20), max(args.warmup, 3)
    for _ in range(warmup):
        enc(x)
    torch.cuda.synchronize()
    lib.crdpn_timing_enable(1)
    tot, n = ctypes.c_double(), ctypes.c_uint64()
    lib.crdpn_timing_read(1, ctypes.byref(tot), ctypes.byref(n))
    l0 = pkg._native.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        enc(x)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    lib.crdpn_timing_read(1, ctypes.byref(tot), ctypes.byref(n))
    lib.crdpn_timing_enable(0)
    launches = (pkg._native.launch_count() - l0) // steps
    kms = tot.value / max(n.value, 1)
    # end to end: pinned host clouds -> device, forward, features back to the host
    for _ in range(3):
        enc(x_host.to(dev, non_blocking=True)).cpu()
    t0 = time.perf_counter()
    for _ in range(steps):
        enc(x_host.to(dev, non_blocking=True)).cpu()
    e2e_ms = (time.perf_counter() - t0) * 1e3 / steps
    tflops = FLOP_PER_POINT * B * P / (kms * 1e-3) / 1e12
    out = {"workload": f"pointnet_eval_B{B}_P{P}_3-64-128-{F}_bf16", "metric": "pointnet_points_per_sec",
           "value": B * P / (ms * 1e-3), "unit": "points/s", "ms_per_step": ms, "kernel_ms": kms,
           "launches_per_step": launches,
           "roofline": {"bound": "tensor", "achieved": tflops, "peak": tf_peak, "unit": "TFLOP/s", "frac": tflops / tf_peak,
                        "peak_kind": peak_kind, "kernel": "pointnet_fwd_eval_kernel"},
           "e2e": {"value": B * P / (e2e_ms * 1e-3), "unit": "points/s", "h2d_bytes_per_step": x_host.numel() * 4,
                   "d2h_bytes_per_step": B * F * 4}}
    if not os.environ.get("CRDPN_BENCH_QUICK"):
        torch.set_num_threads(os.cpu_count() or 1)
        xs = x_host[:16]
        po.forward(xs, st, dtype=torch.float32)
        t0 = time.perf_counter()
        po.forward(xs, st, dtype=torch.float32)
        dt = time.perf_counter() - t0
        out["cpu_baseline"] = {"value": 16 * P / dt, "unit": "points/s", "cores": torch.get_num_threads(), "kind": "port",
                               "sample": f"16 of {B} clouds, fp32 torch ops on CPU (oracle restatement of ShapeEncoderPC)"}
    return out
