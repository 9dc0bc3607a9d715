"""CPU experiment (round 2): which precision recipe brings PointNet train-mode GRADIENTS within 1e-2 of the fp reference?

Recipes, all evaluated with the pinned oracle graph (oracle/pointnet_oracle.py) on the golden case and on a larger one:
  bf16      : round-1 recipe (h1/h2/W2/W3 bf16 everywhere)                                  -> routing flips
  split_fwd : forward values exact up to ~2^-17 relative noise (3-MMA hi/lo split), gates + arg-max from them;
              backward GEMMs on bf16-rounded h2 / W2 / W3 (what pointnet_backward.cu uses)
  split_all : forward split, backward operands hi/lo too (reference arithmetic)
"""
import sys
from pathlib import Path
import numpy as np, torch
ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
from oracle import pointnet_oracle as po

def r(t): return t.detach().to(torch.bfloat16).to(t.dtype)
def ste(t): return t + (r(t) - t.detach())
def split2(t):
    hi = r(t); lo = r(t.detach() - hi)
    return hi, lo

def fwd(x, p, recipe):
    h = x.double()
    for n in (1, 2, 3):
        W = p[f"conv{n}.weight"].double()[:, :, 0]
        if n == 1 or recipe == "fp":
            y = torch.einsum("oc,bcp->bop", W, h)
        elif recipe == "bf16":
            y = torch.einsum("oc,bcp->bop", ste(W), h)
        else:
            Wh, Wl = split2(W); hh, hl = split2(h)
            val = (torch.einsum("oc,bcp->bop", Wh, hh) + torch.einsum("oc,bcp->bop", Wl, hh) + torch.einsum("oc,bcp->bop", Wh, hl)).float().double()
            if recipe == "split_fwd":
                g = torch.einsum("oc,bcp->bop", ste(W), ste(h))     # gradient path: bf16 operands
            else:
                g = torch.einsum("oc,bcp->bop", W, h)
            y = g + (val - g).detach()
        y = y + p[f"conv{n}.bias"].double()[None, :, None]
        y = po._bn(y, p, n, True, None)
        if n < 3:
            h = torch.relu(y)
            if recipe == "bf16":
                h = ste(h)
        else:
            h = y
    return h.max(dim=2).values

def grads(x, st, gout, recipe):
    p = {k: (v.clone().double().requires_grad_() if v.is_floating_point() and "running" not in k else v) for k, v in st.items()}
    out = fwd(x, p, recipe)
    (out * gout).sum().backward()
    return out.detach(), {k: v.grad for k, v in p.items() if getattr(v, "grad", None) is not None}

def run(B, P, F, seed):
    st = po.random_state(F, seed=seed); x = po.random_clouds(B, P, seed=seed + 1)
    gout = torch.randn(B, F, generator=torch.Generator().manual_seed(seed + 2)).double()
    o0, g0 = grads(x, st, gout, "fp")
    for rec in ("bf16", "split_fwd", "split_all"):
        o, g = grads(x, st, gout, rec)
        dev = {k: ((g[k] - g0[k]).norm() / (g0[k].norm() + 1e-30)).item() for k in g0 if not (k.startswith("conv") and k.endswith("bias"))}
        mx = {k: ((g[k] - g0[k]).abs().max() / (g0[k].abs().max() + 1e-30)).item() for k in dev}
        print(f"B={B} P={P} F={F} {rec:10s} out {((o-o0).abs().max()/o0.abs().max()).item():.2e} | " + " ".join(f"{k.replace('.weight','.w').replace('.bias','.b')}:{v:.1e}/{mx[k]:.1e}" for k, v in dev.items()))

if __name__ == "__main__":
    torch.set_num_threads(8)
    run(3, 333, 1024, 46)
    run(8, 2500, 1024, 5)
