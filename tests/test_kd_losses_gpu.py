"""GPU parity of the in-batch NCE KD losses and the KD loss mixer (package kd_losses.py -> crdpn_nce_kd_* / crdpn_kd_mix_*)
against oracle/kd_losses_oracle.py (fp64, pinned to the reference by tests/test_oracle_kd_losses.py) and against the
reference's own fp32 results in tests/golden/kd_losses_golden.npz.

Tolerances (north_star: 1e-4 relative in fp32): losses 1e-4 relative (2e-5 measured), gradients 1e-4 of the tensor's max.
poseNCE_KD's `linear` / `sin` / `sqrt` weightings are compared with the ORACLE at that bar and with the reference's golden
values at the reference's own noise floor (its diagonal weights are fp32 acos noise; see test_oracle_kd_losses.py)."""
from pathlib import Path

import numpy as np
import pytest
import torch

from oracle import kd_losses_oracle as ko

pytestmark = pytest.mark.gpu
GOLD = Path(__file__).parent / "golden" / "kd_losses_golden.npz"


def close(got, ref, tol=1e-4):
    got = got.detach().double().cpu().numpy() if torch.is_tensor(got) else np.asarray(got, dtype=np.float64)
    ref = ref.detach().double().cpu().numpy() if torch.is_tensor(ref) else np.asarray(ref, dtype=np.float64)
    assert got.shape == ref.shape, (got.shape, ref.shape)
    err = np.abs(got - ref).max()
    assert err <= tol * np.abs(ref).max() + 1e-9, (err, np.abs(ref).max())  # 1e-9: exact zeros (e.g. B = 1) vs fp64 dust


def leaf64(t):
    return t.detach().cpu().double().clone().requires_grad_()


def dleaf(t, dev):
    return t.detach().to(dev).clone().requires_grad_()


@pytest.fixture(scope="module")
def gold():
    g = np.load(GOLD)
    out = [torch.from_numpy(g[f"in/out{i}"]) for i in range(6)]
    tout = [torch.from_numpy(g[f"in/tout{i}"]) for i in range(6)]
    return g, out, tout, torch.from_numpy(g["in/sf"]), torch.from_numpy(g["in/tf"]), torch.from_numpy(g["in/label"])


@pytest.mark.parametrize("tau", [0.1, 0.5])
def test_infonce_kd_golden(pkg, cuda, gold, tau):
    g, _, _, sf, tf, label = gold
    pkg.kd_losses.set_dropout_stream(int(g["seed"]), int(g["offset"]))
    a, p = dleaf(sf, cuda), dleaf(tf, cuda)
    loss = pkg.infoNCE_KD(a, p, label.to(cuda), tau)
    (3.0 * loss).backward()
    close(loss, g[f"infonce_kd/tau{tau}/loss"])
    close(a.grad / 3.0, g[f"infonce_kd/tau{tau}/d_ori"])
    close(p.grad / 3.0, g[f"infonce_kd/tau{tau}/d_pos"])
    # the stream advanced by ceil(n*C/4) blocks
    assert pkg.kd_losses._stream_state["offset"] == int(g["offset"]) + (sf.numel() + 3) // 4


@pytest.mark.parametrize("weighting", ["linear", "square", "sqrt", "sin", "sinsin"])
def test_posence_kd_oracle_and_golden(pkg, cuda, gold, weighting):
    g, _, _, sf, tf, label = gold
    a, p = dleaf(sf, cuda), dleaf(tf, cuda)
    loss = pkg.poseNCE_KD(a, p, label.to(cuda), 0.1, weighting)
    loss.backward()
    a64, p64 = leaf64(sf), leaf64(tf)
    ref = ko.nce_kd(a64, p64, label, 0.1, weighting)
    ref.backward()
    # sqrt has an unbounded derivative at 0: pairs of (near-)identical poses amplify fp32 rounding of the distance
    tol = 2e-3 if weighting == "sqrt" else 1e-4
    close(loss, ref, tol)
    close(a.grad, a64.grad, 10 * tol)
    close(p.grad, p64.grad, 10 * tol)
    ltol = {"square": 1e-4, "sinsin": 1e-4, "linear": 5e-4, "sin": 2e-3, "sqrt": 3e-2}[weighting]
    close(loss, g[f"posence_kd/{weighting}/loss"], ltol)


@pytest.mark.parametrize("B,C,tau", [(1, 8, 0.5), (7, 33, 0.07), (138, 200, 0.5), (300, 517, 0.2)])
def test_nce_kd_shapes_against_oracle(pkg, cuda, B, C, tau):
    g = torch.Generator().manual_seed(B * 1000 + C)
    sf = torch.randn(B, C, generator=g)
    tf = torch.randn(B, C, generator=g) + 0.3 * sf
    label = torch.stack((torch.randint(0, 360, (B,), generator=g), torch.randint(0, 180, (B,), generator=g),
                         torch.randint(0, 360, (B,), generator=g)), dim=1)
    if B > 2:
        label[1] = label[0]  # two samples with the same pose: weight exactly/nearly 0 off the diagonal too
    # infoNCE_KD with this build's dropout stream
    pkg.kd_losses.set_dropout_stream(99, 5)
    a, p = dleaf(sf, cuda), dleaf(tf, cuda)
    loss = pkg.infoNCE_KD(a, p, None, tau)
    loss.backward()
    keep = ko.philox_keep_mask(99, 5, B * C, 0.3) if B * C <= 40000 else None
    if keep is not None:
        a64, p64 = leaf64(sf), leaf64(tf)
        ref = ko.nce_kd(a64, p64, None, tau, "none", keep, 0.3)
        ref.backward()
        close(loss, ref)
        close(a.grad, a64.grad)
        close(p.grad, p64.grad)
        assert torch.equal(p.grad.cpu() == 0, torch.from_numpy(~keep).reshape(B, C) | (p.grad.cpu() == 0))
    # poseNCE_KD
    a, p = dleaf(sf, cuda), dleaf(tf, cuda)
    loss = pkg.poseNCE_KD(a, p, label.to(cuda), tau, "sinsin")
    loss.backward()
    a64, p64 = leaf64(sf), leaf64(tf)
    ref = ko.nce_kd(a64, p64, label, tau, "sinsin")
    ref.backward()
    close(loss, ref)
    close(a.grad, a64.grad, 2e-4)
    close(p.grad, p64.grad, 2e-4)


def test_nce_kd_is_deterministic_and_needs_cuda(pkg, cuda):
    sf, tf = torch.randn(46, 200), torch.randn(46, 200)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        pkg.infoNCE_KD(sf, tf, None, 0.5)
    outs = []
    for _ in range(2):
        pkg.kd_losses.set_dropout_stream(7, 0)
        a = dleaf(sf, cuda)
        loss = pkg.infoNCE_KD(a, tf.to(cuda), None, 0.5)
        loss.backward()
        outs.append((loss.item(), a.grad.clone()))
    assert outs[0][0] == outs[1][0] and torch.equal(outs[0][1], outs[1][1])
    with pytest.raises(ValueError):
        pkg.poseNCE_KD(sf.to(cuda), tf.to(cuda), torch.zeros(46, 3, device=cuda), 0.5, "cubic")


@pytest.mark.parametrize("T", [1.0, 2.0])
def test_kl_golden(pkg, cuda, gold, T):
    g, out, tout, *_ = gold
    s, t = dleaf(out[0], cuda), dleaf(tout[0], cuda)
    loss = pkg.TemperatureScaledKLDivLoss(T)(s, t)
    loss.backward()
    close(loss, g[f"kl/T{T}/loss"])
    close(s.grad, g[f"kl/T{T}/d_student"])
    close(t.grad, g[f"kl/T{T}/d_teacher"])


def test_ce_and_delta_golden(pkg, cuda, gold):
    g, out, tout, sf, tf, label = gold
    s = dleaf(out[1], cuda)
    loss = pkg.CELoss(180)(s, label[:, 1].to(cuda))
    loss.backward()
    close(loss, g["ce180/loss"])
    close(s.grad, g["ce180/d_pred"])
    d = [dleaf(out[3], cuda), dleaf(out[4], cuda), dleaf(out[5], cuda)]
    loss = pkg.DeltaLoss(15)(d[0], d[1], d[2], label.to(cuda).float())
    loss.backward()
    close(loss, g["delta/loss"])
    for i in range(3):
        close(d[i].grad, g[f"delta/d_pred{i}"])


def test_student_step_loss_golden_fused_and_piecewise(pkg, cuda, gold):
    g, out, tout, sf, tf, label = gold
    lab = label.to(cuda)
    # (1) everything in one launch
    o, to, a, p = [dleaf(t, cuda) for t in out], [dleaf(t, cuda) for t in tout], dleaf(sf, cuda), dleaf(tf, cuda)
    loss = pkg.student_kd_step_loss(o, to, a, p, lab)
    loss.backward()
    close(loss, g["step/loss"])
    for i in range(6):
        close(o[i].grad, g[f"step/d_out{i}"])
        close(to[i].grad, g[f"step/d_tout{i}"])
    close(a.grad, g["step/d_sf"])
    close(p.grad, g["step/d_tf"])
    # (2) the reference's own call sequence (base_class.py:365-387) with the mirrored classes
    o2, to2, a2, p2 = [dleaf(t, cuda) for t in out], [dleaf(t, cuda) for t in tout], dleaf(sf, cuda), dleaf(tf, cuda)
    gt = (pkg.CELoss(360)(o2[0], lab[:, 0]) + pkg.CELoss(180)(o2[1], lab[:, 1]) + pkg.CELoss(360)(o2[2], lab[:, 2]) +
          pkg.DeltaLoss(15)(o2[3], o2[4], o2[5], lab.float()))
    close(gt, g["step/gt_loss"])
    close(pkg.kd_losses.gt_loss(o2, lab), g["step/gt_loss"])
    loss2 = pkg.calculate_kd_loss_new(o2, to2, a2, p2, gt)
    loss2.backward()
    close(loss2, g["step/loss"])
    for i in range(6):
        close(o2[i].grad, g[f"step/d_out{i}"])
    close(a2.grad, g["step/d_sf"])


@pytest.mark.parametrize("n,C", [(1, 5), (138, 200), (500, 1000)])
def test_step_loss_shapes_against_oracle(pkg, cuda, n, C):
    out, tout, sf, tf, label = ko.synthetic_step(n, C, seed=n + C)
    out[0][0, 3] = 60.0   # a saturated logit: softmax / log-sum-exp stability
    tout[1][0, 2] = -80.0
    o, to, a, p = [dleaf(t, cuda) for t in out], [dleaf(t, cuda) for t in tout], dleaf(sf, cuda), dleaf(tf, cuda)
    loss = pkg.student_kd_step_loss(o, to, a, p, label.to(cuda), temperature=2.0)
    up = torch.tensor(0.37, device=cuda)
    (loss * up).backward()
    o64, to64, a64, p64 = [leaf64(t) for t in out], [leaf64(t) for t in tout], leaf64(sf), leaf64(tf)
    ref = ko.student_kd_step_loss(o64, to64, a64, p64, label, temperature=2.0)
    (ref * 0.37).backward()
    close(loss, ref)
    for i in range(6):
        close(o[i].grad, o64[i].grad)
        close(to[i].grad, to64[i].grad)
    close(a.grad, a64.grad)
    close(p.grad, p64.grad)
    # teacher side detached (the KD loop's frozen teacher): no teacher gradients are produced or needed
    o3 = [dleaf(t, cuda) for t in out]
    loss3 = pkg.student_kd_step_loss(o3, [t.to(cuda) for t in tout], dleaf(sf, cuda), tf.to(cuda), label.to(cuda), temperature=2.0)
    loss3.backward()
    assert loss3.item() == loss.item()
    close(o3[0].grad, o64[0].grad / 0.37)


# ---- the other in-batch variants (auxiliary/model_utils.py:169-223, 288-351) -----------------------------------------
def test_infonce_single_multipose_golden(pkg, cuda, gold):
    """infoNCE / singleinfoNCE_KD / multiposeNCE_KD against the REFERENCE's fp32 values and input gradients."""
    g, _, _, sf, tf, label = gold
    a, p = dleaf(sf, cuda), dleaf(tf, cuda)
    loss = pkg.infoNCE(a, p, 0.1)
    (2.0 * loss).backward()
    close(loss, g["infonce/loss"])
    close(a.grad / 2.0, g["infonce/d_ori"])
    close(p.grad / 2.0, g["infonce/d_pos"])
    a, p = dleaf(sf, cuda), dleaf(tf, cuda)
    loss = pkg.singleinfoNCE_KD(a, p, label.to(cuda), 0.1)
    loss.backward()
    close(loss, g["single/loss"])
    close(a.grad, g["single/d_ori"])
    close(p.grad, g["single/d_pos"])
    a, p = dleaf(sf, cuda), dleaf(tf, cuda)
    loss = pkg.multiposeNCE_KD(a, p, torch.from_numpy(g["multipose/label"]).to(cuda), 0.1)
    loss.backward()
    close(loss, g["multipose/loss"])
    close(a.grad, g["multipose/d_ori"])
    close(p.grad, g["multipose/d_pos"])


@pytest.mark.parametrize("weighting", ["linear", "square", "sqrt", "sin", "sinsin"])
def test_posence_oracle_and_golden(pkg, cuda, gold, weighting):
    """poseNCE against the oracle (1e-4; `sqrt` has an unbounded derivative at distance 0: 2e-3) and, for the weightings
    that square the reference's diagonal noise away, against its golden values."""
    g, _, _, sf, tf, label = gold
    a, p = dleaf(sf, cuda), dleaf(tf, cuda)
    loss = pkg.poseNCE(a, p, label.to(cuda), 0.1, weighting)
    loss.backward()
    a64, p64 = leaf64(sf), leaf64(tf)
    want = ko.nce_self(a64, p64, label, 0.1, weighting)
    want.backward()
    tol = 2e-3 if weighting == "sqrt" else 1e-4
    close(loss, want, tol)
    close(a.grad, a64.grad, tol)
    close(p.grad, p64.grad, tol)
    if weighting == "square":
        close(loss, g["posence/square/loss"], 1e-4)
        close(a.grad, g["posence/square/d_ori"], 3e-4)


@pytest.mark.parametrize("B,C,tau", [(46, 200, 0.1), (138, 200, 0.5), (7, 33, 0.07), (1, 16, 0.1)])
def test_variant_shapes_against_oracle(pkg, cuda, B, C, tau):
    gen = torch.Generator().manual_seed(B * 1000 + C)
    sf, tf = torch.randn(B, C, generator=gen), torch.randn(B, C, generator=gen)
    label = torch.stack([torch.rand(B, generator=gen) * 360, torch.rand(B, generator=gen) * 178 + 1, torch.rand(B, generator=gen) * 360], 1)
    if B > 6:
        label = ko.clustered_labels(label, seed=B)
    cases = [(lambda a, p: pkg.infoNCE(a, p, tau), lambda a, p: ko.nce_self(a, p, None, tau)),
             (lambda a, p: pkg.poseNCE(a, p, label.to(cuda), tau, "sinsin"), lambda a, p: ko.nce_self(a, p, label, tau, "sinsin")),
             (lambda a, p: pkg.singleinfoNCE_KD(a, p, None, tau), lambda a, p: ko.single_nce_kd(a, p, tau)),
             (lambda a, p: pkg.multiposeNCE_KD(a, p, label.to(cuda), tau), lambda a, p: ko.multipose_nce_kd(a, p, label, tau))]
    for got_fn, want_fn in cases:
        a, p = dleaf(sf, cuda), dleaf(tf, cuda)
        got = got_fn(a, p)
        got.backward()
        a64, p64 = leaf64(sf), leaf64(tf)
        want = want_fn(a64, p64)
        want.backward()
        close(got, want)
        close(a.grad, a64.grad, 2e-4)
        close(p.grad, p64.grad, 2e-4)
