#!/usr/bin/env python
"""bench_sweep.py -- BASELINE.json configs[4]: negatives sweep K x D x B on one B200 to map the HBM roofline.

    python bench_sweep.py [--steps 20] [--warmup 3] [--out profiles/rN_sweep.json]

Every point is one `crdpn_crd_step` (score + loss + backward in one pass, then reduction + momentum update) over
N = 1M-row fp32 banks (1.02 GB at D=128, 2.05 GB at D=256: far larger than L2, random rows).  Reported per point:
step time, event-timed score-kernel time, scores/s, algorithmic GB/s (SURVEY.md 8d byte count) and its fraction of
the measured HBM copy peak.  The largest point (B=512, K=131072, D=256) gathers 137 GB per step -- the stock
index_select+bmm formulation would have to materialise 68.7 GB per bank; here nothing of that size exists.
"""
from __future__ import annotations

import argparse
import json
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))
import bench  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--out", default="")
    ap.add_argument("--quick", action="store_true", help="corner points only")
    args = ap.parse_args()
    import torch
    import __graft_entry__ as ge
    pkg = ge.load_package()
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(dev)
    hbm_peak, _, peak_kind = bench.measured_peaks()
    Ks = [4096, 16384, 65536, 131072]
    Ds = [128, 256]
    Bs = [46, 128, 512]
    if args.quick:
        Ks, Bs = [4096, 131072], [46, 512]
    rows = []
    for D in Ds:
        for B in Bs:
            for K in Ks:
                c = dict(B=B, D=D, K=K, N=1_000_000, s_dim=256, t_dim=256, T=0.07, m=0.5)
                work = 2 * B * (K + 1) * D * 4
                steps = max(3, min(args.steps, int(40e9 // work)))  # bound the big points to ~40 GB of gathers
                r = bench.time_crd_resident(pkg, torch, dev, c, steps, args.warmup)
                ms = r["total_ms"] / steps
                ab = bench.algorithmic_bytes(c)
                row = {"B": B, "D": D, "K": K, "N": c["N"], "steps": steps, "ms_per_step": round(ms, 4),
                       "kernel_ms": round(r["kernel_ms_avg"], 4),
                       "scores_per_s": bench.scores_per_step(c) / (ms * 1e-3),
                       "algorithmic_bytes": ab, "achieved_gbs": round(ab / (r["kernel_ms_avg"] * 1e-3) / 1e9, 1),
                       "frac_of_peak": round(ab / (r["kernel_ms_avg"] * 1e-3) / 1e9 / hbm_peak, 3)}
                rows.append(row)
                print(json.dumps(row), flush=True)
                torch.cuda.empty_cache()
    out = {"what": "CRD negatives sweep (BASELINE configs[4]) on 1x B200", "hbm_peak_gbs": hbm_peak,
           "peak_kind": peak_kind, "gpu": torch.cuda.get_device_name(0), "rows": rows}
    if args.out:
        Path(args.out).parent.mkdir(parents=True, exist_ok=True)
        Path(args.out).write_text(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
