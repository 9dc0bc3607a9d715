"""CPU tests pinning the PointNet oracle (oracle/pointnet_oracle.py) to the reference's ShapeEncoderPC
(auxiliary/model.py:154-180): against tests/golden/pointnet_golden.npz (made from the reference by
oracle/gen_golden.py) everywhere, and against the reference module itself where /root/reference exists."""
from pathlib import Path

import numpy as np
import pytest
import torch

from oracle import pointnet_oracle as po

GOLD = Path(__file__).parent / "golden" / "pointnet_golden.npz"


@pytest.fixture(scope="module")
def gold():
    g = np.load(GOLD)
    st = {k[6:]: torch.from_numpy(g[k]) for k in g.files if k.startswith("state/")}
    return g, st


def test_eval_forward_matches_golden(gold):
    g, st = gold
    out = po.forward(torch.from_numpy(g["x"]), st, training=False)
    np.testing.assert_allclose(out.float().numpy(), g["eval_out"], rtol=1e-5, atol=2e-6)


def test_train_forward_stats_and_grads_match_golden(gold):
    g, st = gold
    st = {k: (v.clone().double().requires_grad_() if v.is_floating_point() and "running" not in k else v) for k, v in st.items()}
    ns = {}
    out = po.forward(torch.from_numpy(g["x"]), st, training=True, new_stats=ns)
    np.testing.assert_allclose(out.detach().float().numpy(), g["train_out"], rtol=1e-4, atol=1e-5)
    for k, v in ns.items():
        np.testing.assert_allclose(v.detach().double().numpy(), g["after/" + k].astype(np.float64), rtol=1e-5, atol=1e-6)
    (out * torch.from_numpy(g["gout"]).double()).sum().backward()
    for k in g.files:
        if k.startswith("grad/"):
            ref = g[k].astype(np.float64)
            got = st[k[5:]].grad.numpy()
            if k.startswith("grad/conv") and k.endswith(".bias"):
                # a conv bias feeding train-mode BN has an exactly-zero gradient (the batch mean removes it);
                # the fp32 reference only holds rounding noise there
                wscale = np.abs(g[k.replace(".bias", ".weight")]).max()
                assert np.abs(ref).max() < 1e-3 * wscale and np.abs(got).max() < 1e-9 * wscale, k
                continue
            scale = np.abs(ref).max() + 1e-12
            assert np.abs(got - ref).max() / scale < 2e-4, k


def test_bf16_recipe_within_tolerance(gold):
    g, st = gold
    out = po.forward_bf16_emulated(torch.from_numpy(g["x"]), st)
    ref = torch.from_numpy(g["eval_out"])
    assert ((out - ref).abs().max() / ref.abs().max()).item() < 1e-2   # north_star bf16 tolerance


@pytest.mark.skipif(not Path("/root/reference/auxiliary/model.py").exists(), reason="reference not mounted")
@pytest.mark.parametrize("training", [False, True])
def test_against_live_reference_module(training):
    st = po.random_state(256, seed=11)
    x = po.random_clouds(2, 97, seed=12)
    ref = po.make_reference_module(st, 256, training)
    with torch.no_grad():
        want = ref(x)
    got = po.forward(x, st, training=training).float()
    np.testing.assert_allclose(got.numpy(), want.numpy(), rtol=1e-4, atol=1e-5)


def test_max_commutes_with_bn3_sign_trick(gold):
    """max_p bn3(y) == bn3(max_p y) where scale >= 0 and bn3(min_p y) where scale < 0 (used by the kernel)."""
    g, st = gold
    x = torch.from_numpy(g["x"]).double()
    h = x
    for n in (1, 2):
        W = st[f"conv{n}.weight"].double()[:, :, 0]
        h = torch.relu(po._bn(torch.einsum("oc,bcp->bop", W, h) + st[f"conv{n}.bias"].double()[None, :, None], st, n, False, None))
    y = torch.einsum("oc,bcp->bop", st["conv3.weight"].double()[:, :, 0], h) + st["conv3.bias"].double()[None, :, None]
    scale = st["bn3.weight"].double() * torch.rsqrt(st["bn3.running_var"].double() + po.BN_EPS)
    pick = torch.where(scale[None, :] >= 0, y.max(2).values, y.min(2).values)
    out = (pick - st["bn3.running_mean"].double()) * scale + st["bn3.bias"].double()
    np.testing.assert_allclose(out.numpy(), po.forward(x, st).numpy(), rtol=1e-10, atol=1e-10)
    assert (scale < 0).sum() > 100


def _grads(fn, x, st, gout):
    p = {k: (v.clone().double().requires_grad_() if v.is_floating_point() and "running" not in k else v) for k, v in st.items()}
    out = fn(x, p)
    (out * gout).sum().backward()
    return out.detach(), {k: v.grad for k, v in p.items() if getattr(v, "grad", None) is not None}


def test_bf16_emulated_train_oracle_is_the_oracle_without_rounding(gold, monkeypatch):
    """forward_train_bf16_emulated == forward(training=True) once rounding is switched off: same graph, so it
    inherits the pin to the reference; with rounding on, features stay within the bf16 tolerance."""
    g, st = gold
    x, gout = torch.from_numpy(g["x"]), torch.from_numpy(g["gout"]).double()
    o_ref, g_ref = _grads(lambda a, p: po.forward(a, p, training=True), x, st, gout)
    o_emu, _ = _grads(po.forward_train_bf16_emulated, x, st, gout)
    assert ((o_emu - o_ref).abs().max() / o_ref.abs().max()).item() < 1e-2
    monkeypatch.setattr(po, "_ste_bf16", lambda t: t)
    o_id, g_id = _grads(po.forward_train_bf16_emulated, x, st, gout)
    assert torch.allclose(o_id, o_ref, rtol=1e-12, atol=1e-12)
    for k in g_ref:
        assert torch.allclose(g_id[k], g_ref[k], rtol=1e-9, atol=1e-12), k


def test_bf16_gradient_deviation_is_routing_not_arithmetic(gold):
    """Why train-mode GRADIENTS of a bf16 pipeline cannot meet 1e-2 against the fp32 reference while the features
    do: max-pool and ReLU route gradients discontinuously, and bf16 rounding flips near-tied arg-max points / ReLU
    gates.  Evidence: conv3.weight's gradient differs by >3% between the fp and the bf16-recipe oracle, but by
    <1.5% once the fp oracle is forced to route through the bf16 run's arg-max points."""
    g, st = gold
    x, gout = torch.from_numpy(g["x"]), torch.from_numpy(g["gout"]).double()

    def body(a, p, rounded):
        h = a.double()
        for n in (1, 2, 3):
            W = p[f"conv{n}.weight"].double()[:, :, 0]
            if rounded and n > 1:
                W = po._ste_bf16(W)
            y = torch.einsum("oc,bcp->bop", W, h) + p[f"conv{n}.bias"].double()[None, :, None]
            y = po._bn(y, p, n, True, None)
            h = (po._ste_bf16(torch.relu(y)) if rounded else torch.relu(y)) if n < 3 else y
        return h

    idx = body(x, {k: v for k, v in st.items()}, True).argmax(dim=2)
    _, g_fp = _grads(lambda a, p: po.forward(a, p, training=True), x, st, gout)
    _, g_emu = _grads(po.forward_train_bf16_emulated, x, st, gout)
    _, g_forced = _grads(lambda a, p: body(a, p, False).gather(2, idx[:, :, None])[:, :, 0], x, st, gout)
    nrel = lambda a, b: ((a - b).norm() / b.norm()).item()
    k = "conv3.weight"
    assert nrel(g_emu[k], g_fp[k]) > 3e-2
    assert nrel(g_emu[k], g_forced[k]) < 1.5e-2
