"""GPU parity tests of the train-mode PointNet encoder (batch-statistics BatchNorm forward + backward) against
the oracle pinned to the reference's ShapeEncoderPC (auxiliary/model.py:154-180; training.py:30,47,75).

Default train precision ("fp32": every tensor-core product as three fp16 hi/lo MMAs, csrc/pointnet_train_split.cu):
  * features vs the fp reference / golden vectors: <= 1e-4 (north_star's fp32 bar; measured ~3e-7)
  * running statistics: <= 1e-5
  * EVERY parameter gradient vs the reference's own golden gradients / the fp oracle: <= 1e-2 norm-relative
    (north_star's bf16 bar: the backward GEMMs run on bf16 operands; the ROUTING -- arg-max points, ReLU gates -- is
    the fp32 one, which is what the 12-18 % deviation of round 1 came from)
``train_precision = "bf16"`` (round 1's recipe, one MMA per product, ~1.3x faster):
  * features / statistics <= 1e-2 / 2e-3; gradients vs the exact gradient of the SAME bf16 recipe (oracle
    forward_train_bf16_emulated) <= 1e-2 on average; vs the fp reference they deviate by the routing effect
    (tests/test_oracle_pointnet.py) and are only loosely bounded."""
from pathlib import Path

import numpy as np
import pytest
import torch

from oracle import pointnet_oracle as po

pytestmark = pytest.mark.gpu
GOLD = Path(__file__).parent / "golden" / "pointnet_golden.npz"
TOL = 1e-2


def _nrel(got, want):
    got, want = got.detach().double().cpu(), want.detach().double().cpu()
    return ((got - want).norm() / (want.norm() + 1e-300)).item()


def _maxrel(got, want):
    got, want = got.detach().double().cpu(), want.detach().double().cpu()
    return ((got - want).abs().max() / want.abs().max()).item()


def _encoder(pkg, st, F, dev, precision="fp32"):
    enc = pkg.ShapeEncoderPC(F)
    enc.load_state_dict({k: v.clone() for k, v in st.items()})
    enc.train_precision = precision
    return enc.to(dev).train()


def _oracle(fn, x, st, gout, dtype=torch.float64):
    p = {k: (v.clone().to(dtype).requires_grad_() if v.is_floating_point() and "running" not in k else v.clone())
         for k, v in st.items()}
    ns = {}
    out = fn(x, p, ns)
    (out * gout.to(dtype)).sum().backward()
    return out.detach(), {k: v.grad for k, v in p.items() if getattr(v, "grad", None) is not None}, ns


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
def test_train_golden_vector_from_reference(pkg, cuda, precision):
    g = np.load(GOLD)
    st = {k[6:]: torch.from_numpy(g[k]) for k in g.files if k.startswith("state/")}
    enc = _encoder(pkg, st, 1024, cuda, precision)
    fp32 = precision == "fp32"
    x, gout = torch.from_numpy(g["x"]), torch.from_numpy(g["gout"])
    out = enc(x.to(cuda))
    (out * gout.to(cuda)).sum().backward()
    assert out.shape == (3, 1024) and out.dtype == torch.float32
    assert _maxrel(out, torch.from_numpy(g["train_out"])) < (1e-4 if fp32 else TOL)   # reference module's own train output
    for n in (1, 2, 3):
        bn = getattr(enc, f"bn{n}")
        tol = 1e-5 if (n == 1 or fp32) else 2e-3   # BN1 statistics are analytic in fp64; BN2/BN3 come from the MMA pipeline
        assert _nrel(bn.running_mean, torch.from_numpy(g[f"after/bn{n}.running_mean"])) < tol
        assert _nrel(bn.running_var, torch.from_numpy(g[f"after/bn{n}.running_var"])) < tol
        assert int(bn.num_batches_tracked) == int(g[f"after/bn{n}.num_batches_tracked"]) == 1
    # gradients vs the reference's own golden gradients.  fp32 precision: every tensor within north_star's 1e-2.
    # bf16 recipe: BN3's are smooth in the features -> tight; the rest is routed through arg-max points / ReLU gates
    # that bf16 rounding flips (see module docstring) -> loose bound, values printed
    dev = {}
    for name, prm in enc.named_parameters():
        ref = torch.from_numpy(g["grad/" + name])
        if "conv" in name and name.endswith("bias"):
            assert prm.grad.abs().max().item() == 0.0   # exactly zero: the batch mean removes the bias
            continue
        dev[name] = _nrel(prm.grad, ref)
    print(f"gradient deviation vs fp32 reference ({precision}):", {k: round(v, 5) for k, v in dev.items()})
    assert dev["bn3.bias"] < 1e-5 and dev["bn3.weight"] < TOL
    assert max(dev.values()) < (TOL if fp32 else 0.3), dev


@pytest.mark.parametrize("B,P,F", [(1, 64, 128), (2, 50, 128), (3, 333, 1024), (2, 128, 256), (5, 257, 512),
                                    (4, 1000, 256), (149, 40, 128), (8, 2500, 1024)])
def test_train_forward_backward_vs_fp_oracle(pkg, cuda, B, P, F):
    """Default (fp32-accurate) train path against the fp oracle pinned to the reference: features 1e-4, running
    statistics 1e-5, EVERY parameter gradient 1e-2."""
    st = po.random_state(F, seed=B * 13 + P)
    x = po.random_clouds(B, P, seed=P + 1)
    gout = torch.randn(B, F, generator=torch.Generator().manual_seed(F + B))
    dtype = torch.float64 if B * P * F < 4e6 else torch.float32
    want, gwant, ns = _oracle(lambda a, p, ns: po.forward(a, p, training=True, new_stats=ns, dtype=dtype), x, st, gout, dtype)
    enc = _encoder(pkg, st, F, cuda)
    out = enc(x.to(cuda))
    (out * gout.to(cuda)).sum().backward()
    assert out.shape == (B, F)
    assert _maxrel(out, want) < 1e-4
    for n in (1, 2, 3):
        bn = getattr(enc, f"bn{n}")
        assert _nrel(bn.running_mean, ns[f"bn{n}.running_mean"]) < 1e-5
        assert _nrel(bn.running_var, ns[f"bn{n}.running_var"]) < 1e-5
    errs = {}
    for name, prm in enc.named_parameters():
        if "conv" in name and name.endswith("bias"):
            assert prm.grad.abs().max().item() == 0.0
            continue
        errs[name] = _nrel(prm.grad, gwant[name])
    print("gradient deviation vs fp oracle:", {k: round(v, 5) for k, v in errs.items()})
    assert max(errs.values()) < TOL, errs


@pytest.mark.parametrize("B,P,F", [(1, 64, 128), (3, 333, 1024), (5, 257, 512), (149, 40, 128), (8, 2500, 1024)])
def test_train_forward_backward_vs_bf16_recipe_oracle(pkg, cuda, B, P, F):
    st = po.random_state(F, seed=B * 13 + P)
    x = po.random_clouds(B, P, seed=P + 1)
    gout = torch.randn(B, F, generator=torch.Generator().manual_seed(F + B))
    dtype = torch.float64 if B * P * F < 4e6 else torch.float32
    want, gwant, ns = _oracle(lambda a, p, ns: po.forward(a, p, training=True, new_stats=ns, dtype=dtype), x, st, gout, dtype)
    ewant, gemu, _ = _oracle(lambda a, p, ns: po.forward_train_bf16_emulated(a, p, dtype=dtype), x, st, gout, dtype)
    enc = _encoder(pkg, st, F, cuda, "bf16")
    out = enc(x.to(cuda))
    (out * gout.to(cuda)).sum().backward()
    assert out.shape == (B, F)
    assert _maxrel(out, want) < TOL                       # features vs the fp oracle (pinned to the reference)
    assert _nrel(out, ewant) < 2e-3                       # and vs the same precision recipe
    for n in (1, 2, 3):
        bn = getattr(enc, f"bn{n}")
        assert _nrel(bn.running_mean, ns[f"bn{n}.running_mean"]) < 2e-3
        assert _nrel(bn.running_var, ns[f"bn{n}.running_var"]) < 2e-3
    errs = {}
    for name, prm in enc.named_parameters():
        if "conv" in name and name.endswith("bias"):
            assert prm.grad.abs().max().item() == 0.0
            continue
        errs[name] = _nrel(prm.grad, gemu[name])
    # 1e-2 on average; a single tensor may reach 2e-2 when one or two of the B*F arg-max decisions differ between
    # the fp64 emulation and the fp32-accumulating tensor-core pipeline (a discrete re-routing, not arithmetic error)
    assert sum(errs.values()) / len(errs) < TOL, errs
    assert max(errs.values()) < 2 * TOL, errs


def test_two_train_steps_then_eval_use_the_updated_statistics(pkg, cuda):
    st = po.random_state(256, seed=21)
    xs = [po.random_clouds(3, 200, seed=s) for s in (22, 23)]
    enc = _encoder(pkg, st, 256, cuda)
    cur = {k: v.clone() for k, v in st.items()}
    for x in xs:
        with torch.no_grad():
            enc(x.to(cuda))
        ns = {}
        po.forward(x, cur, training=True, new_stats=ns)
        cur.update({k: v.float() if v.is_floating_point() else v for k, v in ns.items()})
    for n in (1, 2, 3):
        bn = getattr(enc, f"bn{n}")
        assert int(bn.num_batches_tracked) == 2
        assert _nrel(bn.running_mean, cur[f"bn{n}.running_mean"]) < 2e-3
        assert _nrel(bn.running_var, cur[f"bn{n}.running_var"]) < 2e-3
    enc.eval()
    x = po.random_clouds(2, 300, seed=24)
    with torch.no_grad():
        out = enc(x.to(cuda))
    sd = {k: v.detach().cpu() for k, v in enc.state_dict().items()}
    assert _maxrel(out, po.forward(x, sd, training=False)) < TOL


def test_train_mode_contract(pkg, cuda):
    st = po.random_state(128, seed=31)
    enc = _encoder(pkg, st, 128, cuda)
    x = po.random_clouds(2, 64, seed=32)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        enc(x)
    with pytest.raises(RuntimeError, match="input point cloud"):
        enc(x.to(cuda).requires_grad_())
    out = enc(x.to(cuda))
    assert out.requires_grad and out.grad_fn is not None
    opt = torch.optim.Adam(enc.parameters(), lr=1e-3)     # training.py:269: Adam over .parameters()
    out.square().mean().backward()
    opt.step()
    assert all(p.grad is not None and torch.isfinite(p.grad).all() for p in enc.parameters())
    # point order inside a cloud does not matter (statistics and max are permutation invariant)
    enc2 = _encoder(pkg, st, 128, cuda)
    enc3 = _encoder(pkg, st, 128, cuda)
    perm = torch.randperm(64)
    a = enc2(x.to(cuda))
    b = enc3(x[:, :, perm].contiguous().to(cuda))
    assert _maxrel(a, b) < 2e-3


def _torch_reference_step(st, x, gout):
    """The reference's own ops (nn.Conv1d + nn.BatchNorm1d in train mode + max, auxiliary/model.py:174-180) in fp32 on
    the host, whole batch at once; returns (out, grads, module)."""
    import torch.nn as nn

    class Enc(nn.Module):
        def __init__(self, F):
            super().__init__()
            self.conv1, self.conv2, self.conv3 = nn.Conv1d(3, 64, 1), nn.Conv1d(64, 128, 1), nn.Conv1d(128, F, 1)
            self.bn1, self.bn2, self.bn3 = nn.BatchNorm1d(64), nn.BatchNorm1d(128), nn.BatchNorm1d(F)

        def forward(self, s):
            s = torch.relu(self.bn1(self.conv1(s)))
            s = torch.relu(self.bn2(self.conv2(s)))
            s = self.bn3(self.conv3(s))
            return torch.max(s, 2, keepdim=True)[0].view(s.shape[0], -1)

    m = Enc(st["conv3.weight"].shape[0]).train()
    m.load_state_dict({k: v.clone() for k, v in st.items()})
    out = m(x)
    (out * gout).sum().backward()
    return out.detach(), {k: p.grad for k, p in m.named_parameters()}, m


def test_config2_full_size_train_step(pkg, cuda):
    """BASELINE configs[1] in train mode at FULL size (batch 160, 2500 points, 1024 features): features, running
    statistics and EVERY parameter gradient against the same layer stack as the reference (nn.Conv1d / nn.BatchNorm1d /
    max, fp32, host) run over the whole batch."""
    st = po.random_state(1024, seed=46)
    x = po.random_clouds(160, 2500, seed=46)
    gout = torch.randn(160, 1024, generator=torch.Generator().manual_seed(7))
    enc = _encoder(pkg, st, 1024, cuda)
    out = enc(x.to(cuda))
    (out * gout.to(cuda)).sum().backward()
    torch.cuda.synchronize()
    assert torch.isfinite(out).all()
    assert all(torch.isfinite(p.grad).all() for p in enc.parameters())
    # BN1 statistics have a closed form in the input moments: check them at full size in fp64
    xd = x.double().permute(1, 0, 2).reshape(3, -1)
    W = st["conv1.weight"].double()[:, :, 0]
    y = W @ xd + st["conv1.bias"].double()[:, None]
    mean, var = y.mean(1), y.var(1, unbiased=True)
    assert _nrel(enc.bn1.running_mean, 0.9 * st["bn1.running_mean"].double() + 0.1 * mean) < 1e-5
    assert _nrel(enc.bn1.running_var, 0.9 * st["bn1.running_var"].double() + 0.1 * var) < 1e-5
    # bn3.bias gradient = column sums of grad_out (exact)
    assert _nrel(enc.bn3.bias.grad, gout.sum(0)) < 1e-5
    torch.set_num_threads(max(torch.get_num_threads(), 8))
    want, gwant, ref = _torch_reference_step(st, x, gout)
    assert _maxrel(out, want) < 1e-4
    for n in (1, 2, 3):
        bn, rb = getattr(enc, f"bn{n}"), getattr(ref, f"bn{n}")
        assert _nrel(bn.running_mean, rb.running_mean) < 1e-4
        assert _nrel(bn.running_var, rb.running_var) < 1e-4
    errs = {}
    for name, prm in enc.named_parameters():
        if "conv" in name and name.endswith("bias"):
            continue
        errs[name] = _nrel(prm.grad, gwant[name])
    print("full-size gradient deviation vs the fp32 layer stack:", {k: round(v, 5) for k, v in errs.items()})
    assert max(errs.values()) < TOL, errs
