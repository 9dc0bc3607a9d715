"""CPU check for next round's fp32-bank streaming kernel: are three bf16 MMAs per product (hi*hi + lo*hi + hi*lo, fp32
accumulation) accurate enough for north_star's 1e-4?  Emulates the score and gradient GEMMs of the CRD step at
B=46, D=128, K=16384 (one bank direction) with rows / embeddings / coefficients split into bf16 hi + lo parts, against fp64.
Usage: python profiles/experiments/bf16_split_accuracy.py"""
import json
import numpy as np
import torch

def bf16(x):
    return torch.from_numpy(np.ascontiguousarray(x, dtype=np.float32)).to(torch.bfloat16).to(torch.float32).numpy()

def split(x):
    hi = bf16(x)
    return hi, bf16(x - hi)

rng = np.random.default_rng(46)
B, D, K, N, T = 46, 128, 16384, 90000, 0.07
bank = rng.uniform(-1, 1, (N, D)).astype(np.float32)
bank /= np.linalg.norm(bank, axis=1, keepdims=True)
v = rng.standard_normal((B, D)).astype(np.float32)
v /= np.linalg.norm(v, axis=1, keepdims=True)
idx = rng.integers(0, N, (B, K + 1))
out = {}
Z = None
res = {}
for mode in ("fp64", "bf16_1", "bf16_3"):
    loss = 0.0
    G = np.zeros((B, D))
    for b in range(B):
        W = bank[idx[b]]
        if mode == "fp64":
            s = W.astype(np.float64) @ v[b].astype(np.float64)
        else:
            wh, wl = split(W)
            vh, vl = split(v[b])
            s = (wh @ vh).astype(np.float32)
            if mode == "bf16_3":
                s = s + (wl @ vh).astype(np.float32) + (wh @ vl).astype(np.float32)
            s = s.astype(np.float64)
        e = np.exp(s / T)
        if Z is None:
            Z = e.mean() * N
        o = e / Z
        mPn = K / N
        c = mPn + 1e-7
        loss += -(np.log(o[0] / (o[0] + c)) + np.log(mPn / (o[1:] + c)).sum()) / B
        d = o / (o + c) / (B * T)
        d[0] = -c / (o[0] + c) / (B * T)
        if mode == "fp64":
            G[b] = d @ W.astype(np.float64)
        else:
            dh, dl = split(d.astype(np.float32))
            g = (dh @ wh).astype(np.float32)
            if mode == "bf16_3":
                g = g + (dl @ wh).astype(np.float32) + (dh @ wl).astype(np.float32)
            G[b] = g
    res[mode] = (loss, G)
ref_l, ref_G = res["fp64"]
for mode in ("bf16_1", "bf16_3"):
    l, G = res[mode]
    out[mode] = {"loss_rel": abs(l - ref_l) / abs(ref_l), "grad_rel_max": float(np.abs(G - ref_G).max() / np.abs(ref_G).max()),
                 "grad_rel_fro": float(np.linalg.norm(G - ref_G) / np.linalg.norm(ref_G))}
print(json.dumps(out, indent=1))
