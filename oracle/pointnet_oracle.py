"""oracle/pointnet_oracle.py -- CPU oracle for the teacher's PointNet encoder.  TEST INFRASTRUCTURE ONLY.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference``
legs may import this module.  The product package never does.

Restates ``ShapeEncoderPC`` (``/root/reference/auxiliary/model.py:154-180``):

    x[B,3,P] -> relu(bn1(conv1)) [B,64,P] -> relu(bn2(conv2)) [B,128,P] -> bn3(conv3) [B,F,P] -> max over P -> [B,F]

with ``Conv1d(k=1)`` written as a per-point matrix product and ``BatchNorm1d`` (eps 1e-5, momentum 0.1,
biased variance for normalisation, unbiased for ``running_var``) written out explicitly, so that the file
travels to the GPU box where ``/root/reference`` does not exist.  PINNED: ``tests/test_oracle_pointnet.py``
checks this restatement against the reference module itself (imported from ``/root/reference`` when present)
and against ``tests/golden/pointnet_golden.npz`` made by ``oracle/gen_golden.py`` from the reference.
"""
from __future__ import annotations

import sys
import types
from pathlib import Path

import torch

PARAM_SHAPES = {  # auxiliary/model.py:162-172, feature_dim = F
    "conv1.weight": (64, 3, 1), "conv1.bias": (64,),
    "conv2.weight": (128, 64, 1), "conv2.bias": (128,),
    "conv3.weight": ("F", 128, 1), "conv3.bias": ("F",),
    "bn1.weight": (64,), "bn1.bias": (64,), "bn1.running_mean": (64,), "bn1.running_var": (64,),
    "bn2.weight": (128,), "bn2.bias": (128,), "bn2.running_mean": (128,), "bn2.running_var": (128,),
    "bn3.weight": ("F",), "bn3.bias": ("F",), "bn3.running_mean": ("F",), "bn3.running_var": ("F",),
}
BN_EPS = 1e-5
BN_MOMENTUM = 0.1


def load_reference(root: str = "/root/reference"):
    """Import the reference's own ShapeEncoderPC (only possible where /root/reference is mounted).

    ``auxiliary/utils.py:4`` imports matplotlib, which is absent: stub it."""
    if not Path(root).exists():
        return None
    for name in ("matplotlib", "matplotlib.pyplot"):
        if name not in sys.modules:
            sys.modules[name] = types.ModuleType(name)
    sys.modules["matplotlib"].use = lambda *a, **k: None
    sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]
    if root not in sys.path:
        sys.path.insert(0, root)
    from auxiliary.model import ShapeEncoderPC  # type: ignore
    return ShapeEncoderPC


def random_state(feature_dim: int = 1024, seed: int = 46, dtype=torch.float32) -> dict:
    """PyTorch-default-like Conv1d init plus randomised BN affine / running stats:
    gamma~N(0,1) incl. negatives, beta~N(0,1), mean~N(0,1), var~U(0.5,2), so folding bugs are visible."""
    g = torch.Generator().manual_seed(seed)
    st = {}
    for cin, cout, n in ((3, 64, 1), (64, 128, 2), (128, feature_dim, 3)):
        bound = 1.0 / (cin ** 0.5)
        st[f"conv{n}.weight"] = ((torch.rand(cout, cin, 1, generator=g) * 2 - 1) * bound).to(dtype)
        st[f"conv{n}.bias"] = ((torch.rand(cout, generator=g) * 2 - 1) * bound).to(dtype)
        st[f"bn{n}.weight"] = torch.randn(cout, generator=g).to(dtype)
        st[f"bn{n}.bias"] = torch.randn(cout, generator=g).to(dtype)
        st[f"bn{n}.running_mean"] = (torch.randn(cout, generator=g) * 0.2).to(dtype)
        st[f"bn{n}.running_var"] = (torch.rand(cout, generator=g) * 1.5 + 0.5).to(dtype)
        st[f"bn{n}.num_batches_tracked"] = torch.tensor(0, dtype=torch.long)
    return st


def random_clouds(B: int, P: int, seed: int = 46, dtype=torch.float32) -> torch.Tensor:
    """Synthetic clouds in [0,1], per-sample global min/max normalised like auxiliary/dataset.py:147-148."""
    g = torch.Generator().manual_seed(seed)
    x = torch.rand(B, 3, P, generator=g, dtype=torch.float64)
    x = x - x.amin(dim=(1, 2), keepdim=True)
    x = x / x.amax(dim=(1, 2), keepdim=True)
    return x.to(dtype)


def _bn(y, st, n, training, new_stats):
    """BatchNorm1d over [B,C,P] (statistics over B and P per channel)."""
    w, b = st[f"bn{n}.weight"].to(y.dtype), st[f"bn{n}.bias"].to(y.dtype)
    if training:
        cnt = y.shape[0] * y.shape[2]
        mean = y.mean(dim=(0, 2))
        var = ((y - mean[None, :, None]) ** 2).mean(dim=(0, 2))  # biased
        if new_stats is not None:
            rm, rv = st[f"bn{n}.running_mean"].to(y.dtype), st[f"bn{n}.running_var"].to(y.dtype)
            new_stats[f"bn{n}.running_mean"] = ((1 - BN_MOMENTUM) * rm + BN_MOMENTUM * mean).detach()
            new_stats[f"bn{n}.running_var"] = ((1 - BN_MOMENTUM) * rv + BN_MOMENTUM * var * cnt / max(cnt - 1, 1)).detach()
            new_stats[f"bn{n}.num_batches_tracked"] = st[f"bn{n}.num_batches_tracked"] + 1
    else:
        mean, var = st[f"bn{n}.running_mean"].to(y.dtype), st[f"bn{n}.running_var"].to(y.dtype)
    inv = torch.rsqrt(var + BN_EPS)
    return (y - mean[None, :, None]) * (inv * w)[None, :, None] + b[None, :, None]


def forward(x: torch.Tensor, st: dict, training: bool = False, new_stats: dict | None = None,
            dtype=torch.float64) -> torch.Tensor:
    """Restated ShapeEncoderPC.forward (model.py:174-180). x [B,3,P] -> [B,F].  Differentiable torch ops."""
    h = x.to(dtype)
    for n in (1, 2, 3):
        W = st[f"conv{n}.weight"].to(dtype)[:, :, 0]
        h = torch.einsum("oc,bcp->bop", W, h) + st[f"conv{n}.bias"].to(dtype)[None, :, None]
        h = _bn(h, st, n, training, new_stats)
        if n < 3:
            h = torch.relu(h)
    return h.max(dim=2).values


def forward_bf16_emulated(x: torch.Tensor, st: dict) -> torch.Tensor:
    """Eval-mode forward with the kernel's precision recipe: BN folded in fp32; layer 1 in fp32;
    h1, h2, W2', W3' rounded to bf16; fp32 accumulate.  Used to separate bf16 rounding from real bugs."""
    f32 = torch.float32

    def fold(n):
        inv = torch.rsqrt(st[f"bn{n}.running_var"].to(f32) + BN_EPS) * st[f"bn{n}.weight"].to(f32)
        W = st[f"conv{n}.weight"].to(f32)[:, :, 0] * inv[:, None]
        b = (st[f"conv{n}.bias"].to(f32) - st[f"bn{n}.running_mean"].to(f32)) * inv + st[f"bn{n}.bias"].to(f32)
        return W, b

    W1, b1 = fold(1)
    W2, b2 = fold(2)
    W3, b3 = fold(3)
    r = lambda t: t.to(torch.bfloat16).to(f32)
    h1 = r(torch.relu(torch.einsum("oc,bcp->bop", W1, x.to(f32)) + b1[None, :, None]))
    h2 = r(torch.relu(torch.einsum("oc,bcp->bop", r(W2).double(), h1.double()).to(f32) + b2[None, :, None]))
    y = torch.einsum("oc,bcp->bop", r(W3).double(), h2.double()).to(f32)
    return y.max(dim=2).values + b3[None, :]


def make_reference_module(st: dict, feature_dim: int, training: bool):
    """Instantiate the reference's ShapeEncoderPC and load `st` into it (None if reference is absent)."""
    cls = load_reference()
    if cls is None:
        return None
    m = cls(feature_dim)
    m.load_state_dict({k: v.clone() for k, v in st.items()})
    m.train(training)
    return m


def _ste_bf16(t: torch.Tensor) -> torch.Tensor:
    """Round to bf16 in the forward, identity in the backward (straight-through)."""
    return t + (t.detach().to(torch.bfloat16).to(t.dtype) - t.detach())


def forward_train_bf16_emulated(x: torch.Tensor, st: dict, dtype=torch.float64) -> torch.Tensor:
    """Train-mode forward with the train kernels' precision recipe (differentiable, straight-through rounding):
    layer 1 in full precision, h1 / h2 / W2 / W3 rounded to bf16, batch statistics taken from the rounded
    pipeline exactly where the kernels take them.  Separates bf16 effects (above all arg-max flips between
    near-tied points, which re-route gradients discontinuously) from real bugs."""
    h = x.to(dtype)
    for n in (1, 2, 3):
        W = st[f"conv{n}.weight"].to(dtype)[:, :, 0]
        if n > 1:
            W = _ste_bf16(W)
        y = torch.einsum("oc,bcp->bop", W, h) + st[f"conv{n}.bias"].to(dtype)[None, :, None]
        y = _bn(y, st, n, True, None)
        h = _ste_bf16(torch.relu(y)) if n < 3 else y
    return h.max(dim=2).values
