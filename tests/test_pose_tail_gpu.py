"""GPU parity of the pose-tail chain kernel (csrc/pose_tail.cu) through FrozenPoseTail / PoseTail against the reference's
own outputs (tests/golden/pose_tail_golden.npz, made by /root/reference's PoseEstimator modules) and the fp64 oracle.
fp32-accurate mode (three bf16 hi/lo MMAs per product): <= 1e-4 of each tensor's max (north_star's fp32 tolerance; measured
~1e-5); bf16 mode: <= 2e-2."""
from pathlib import Path

import numpy as np
import pytest
import torch

from oracle import pose_tail_oracle as pto

pytestmark = pytest.mark.gpu
GOLD = Path(__file__).parent / "golden" / "pose_tail_golden.npz"
TOL = 1e-4


def _rel(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return np.abs(a - b).max() / (np.abs(b).max() + 1e-30)


def _gold():
    g = np.load(GOLD)
    return g, {k[6:]: torch.from_numpy(g[k]) for k in g.files if k.startswith("state/")}


def test_matches_reference_golden(pkg, cuda):
    g, sd = _gold()
    tail = pkg.FrozenPoseTail.from_state_dict(sd).to(cuda)
    sf, img = torch.from_numpy(g["in/shape_feature"]).to(cuda), torch.from_numpy(g["in/img_feature"]).to(cuda)
    n0 = pkg._native.launch_count()
    for it in range(3):      # repeated calls reuse the workspace: the hand-over counters must come back to zero
        outs, x, p = tail(sf, img)
        assert _rel(x.cpu().numpy(), g["out/x"]) < TOL and _rel(p.cpu().numpy(), g["out/projector"]) < TOL
        for i, o in enumerate(outs):
            assert o.shape == g[f"out/head{i}"].shape and _rel(o.cpu().numpy(), g[f"out/head{i}"]) < TOL
    assert pkg._native.launch_count() - n0 == 8 + 3   # eight weight packs once, then ONE launch per call


def _random_reference_size_state(seed=0):
    torch.manual_seed(seed)
    sd = {}
    C = 2048
    for n, (i, o) in enumerate(((C, C), (C, C // 2), (C // 2, C // 4), (C // 4, 200)), 1):
        sd[f"deformNet.conv{n}.weight"] = torch.randn(o, i, 1) / i ** 0.5
        sd[f"deformNet.conv{n}.bias"] = torch.randn(o) * 0.1
        if n < 4:
            sd.update({f"deformNet.bn{n}.weight": torch.randn(o), f"deformNet.bn{n}.bias": torch.randn(o),
                       f"deformNet.bn{n}.running_mean": torch.randn(o) * 0.2, f"deformNet.bn{n}.running_var": torch.rand(o) + 0.5})
    for h, w in zip(pto.HEADS, (24, 12, 24, 24, 12, 24)):
        sd[h + ".weight"], sd[h + ".bias"] = torch.randn(w, 200) / 14, torch.randn(w) * 0.1
    for lin, (i, o) in zip((0, 3, 6), ((1024, 800), (800, 400), (400, 200))):
        sd[f"projector.{lin}.weight"], sd[f"projector.{lin}.bias"] = torch.randn(o, i) / i ** 0.5, torch.randn(o) * 0.1
    for bn, o in ((1, 800), (4, 400)):
        sd.update({f"projector.{bn}.weight": torch.randn(o), f"projector.{bn}.bias": torch.randn(o),
                   f"projector.{bn}.running_mean": torch.randn(o) * 0.2, f"projector.{bn}.running_var": torch.rand(o) + 0.5})
    return sd


def test_reference_size_batches_and_modes(pkg, cuda):
    """The KD-time shapes (138 rows, 1024 + 1024 features) and the training batch of 160 with random weights vs the oracle;
    a ragged small batch; more rows than one call takes (chunked); bf16 mode; run-to-run bit reproducibility."""
    sd = _random_reference_size_state()
    tail = pkg.FrozenPoseTail.from_state_dict(sd).to(cuda)
    for B, seed in ((138, 1), (138, 2), (46, 3), (160, 4), (5, 5), (300, 6)):
        g = torch.Generator().manual_seed(seed)
        sf, img = torch.randn(B, 1024, generator=g), torch.randn(B, 1024, generator=g)
        outs, x, p = tail(sf.to(cuda), img.to(cuda))
        w_outs, w_x, w_p = pto.forward(sd, sf, img)
        assert _rel(x.cpu().numpy(), w_x.numpy()) < TOL and _rel(p.cpu().numpy(), w_p.numpy()) < TOL, B
        assert all(_rel(a.cpu().numpy(), b.numpy()) < TOL for a, b in zip(outs, w_outs)), B
        first = [t.clone() for t in list(outs) + [x, p]]
        outs2, x2, p2 = tail(sf.to(cuda), img.to(cuda))
        assert all(torch.equal(a, b) for a, b in zip(first, list(outs2) + [x2, p2])), "not bit-reproducible"
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        tail(torch.zeros(2, 1024), torch.zeros(2, 1024))
    tail16 = pkg.FrozenPoseTail.from_state_dict(sd, dtype=torch.bfloat16).to(cuda)
    g = torch.Generator().manual_seed(9)
    sf, img = torch.randn(138, 1024, generator=g), torch.randn(138, 1024, generator=g)
    outs, x, p = tail16(sf.to(cuda), img.to(cuda))
    w_outs, w_x, w_p = pto.forward(sd, sf, img)
    assert x.dtype == torch.float32
    assert _rel(x.cpu().numpy(), w_x.numpy()) < 2e-2 and _rel(p.cpu().numpy(), w_p.numpy()) < 2e-2
    assert all(_rel(a.cpu().numpy(), b.numpy()) < 2e-2 for a, b in zip(outs, w_outs))


def test_trainable_tail_loads_reference_checkpoint_and_matches_golden(pkg, cuda):
    """PoseTail carries the reference's parameter names: the golden state dict loads strictly; eval output equals the
    reference's; train mode (training.py:30,47,75) reproduces the reference's outputs, running statistics and EVERY gradient."""
    g, sd = _gold()
    tail = pkg.PoseTail(img_feature_dim=64, shape_feature_dim=32)
    full = dict(sd)
    for k, v in tail.state_dict().items():
        if k.endswith("num_batches_tracked"):
            full.setdefault(k, torch.zeros((), dtype=torch.long))
    tail.load_state_dict(full, strict=True)
    tail = tail.to(cuda)
    sf = torch.from_numpy(g["in/shape_feature"]).to(cuda).requires_grad_(True)
    img = torch.from_numpy(g["in/img_feature"]).to(cuda).requires_grad_(True)
    tail.eval()
    outs, x, p = tail(sf, img)
    assert _rel(x.cpu().numpy(), g["out/x"]) < TOL and _rel(p.cpu().numpy(), g["out/projector"]) < TOL
    tail.train()
    outs, x, p = tail(sf, img)
    assert _rel(x.detach().cpu().numpy(), g["train/out/x"]) < TOL and _rel(p.detach().cpu().numpy(), g["train/out/projector"]) < TOL
    for i, o in enumerate(outs):
        assert _rel(o.detach().cpu().numpy(), g[f"train/out/head{i}"]) < TOL
    loss = sum((o * torch.from_numpy(g[f"train/gin/head{i}"]).to(cuda)).sum() for i, o in enumerate(outs))
    loss = loss + (x * torch.from_numpy(g["train/gin/x"]).to(cuda)).sum() + (p * torch.from_numpy(g["train/gin/projector"]).to(cuda)).sum()
    loss.backward()
    msd = tail.state_dict()
    for k in g.files:
        if k.startswith("train/state/") and "running" in k:
            assert _rel(msd[k[len("train/state/"):]].cpu().numpy(), g[k]) < TOL, k
        if k == "train/state/deformNet.bn1.num_batches_tracked":
            assert int(msd["deformNet.bn1.num_batches_tracked"]) == int(g[k])
    grads = {n: q.grad for n, q in tail.named_parameters()}
    grads["in/shape_feature"], grads["in/img_feature"] = sf.grad, img.grad
    checked = 0
    for k in g.files:
        if not k.startswith("train/grad/"):
            continue
        name = k[len("train/grad/"):]
        want, got = g[k], grads[name].detach().cpu().numpy()
        assert got.shape == want.shape, name
        if np.abs(want).max() < 1e-4:   # biases in front of train-mode BatchNorm: exactly zero up to rounding, both sides
            assert np.abs(got).max() < 1e-3, name
        else:
            assert _rel(got, want) < 5e-4, name
        checked += 1
    assert checked == 38


def test_train_mode_reference_size_vs_oracle(pkg, cuda):
    """Batch 160 x (1024 + 1024): train-mode outputs, saved running statistics and a sample of gradients vs the fp64 oracle."""
    sd = _random_reference_size_state(seed=3)
    tail = pkg.PoseTail(img_feature_dim=1024, shape_feature_dim=1024)
    full = dict(sd)
    for k in tail.state_dict():
        if k.endswith("num_batches_tracked"):
            full[k] = torch.zeros((), dtype=torch.long)
    tail.load_state_dict(full)
    tail = tail.to(cuda).train()
    gen = torch.Generator().manual_seed(11)
    B = 160
    sf, img = torch.randn(B, 1024, generator=gen), torch.randn(B, 1024, generator=gen)
    g_outs = [torch.randn(B, w, generator=gen) for w in (24, 12, 24, 24, 12, 24)]
    g_x, g_p = torch.randn(B, 200, generator=gen), torch.randn(B, 200, generator=gen)
    sfd, imgd = sf.to(cuda).requires_grad_(True), img.to(cuda).requires_grad_(True)
    outs, x, p = tail(sfd, imgd)
    loss = sum((o * gg.to(cuda)).sum() for o, gg in zip(outs, g_outs)) + (x * g_x.to(cuda)).sum() + (p * g_p.to(cuda)).sum()
    loss.backward()
    w_outs, w_x, w_p, new_running, w_grads = pto.train_step_with_grads(sd, sf, img, g_outs, g_x, g_p)
    assert _rel(x.detach().cpu().numpy(), w_x.numpy()) < TOL and _rel(p.detach().cpu().numpy(), w_p.numpy()) < TOL
    assert all(_rel(a.detach().cpu().numpy(), b.numpy()) < TOL for a, b in zip(outs, w_outs))
    msd = tail.state_dict()
    for k, v in new_running.items():
        assert _rel(msd[k].cpu().numpy(), v.numpy()) < TOL, k
    got = {n: q.grad for n, q in tail.named_parameters()}
    got["in/shape_feature"], got["in/img_feature"] = sfd.grad, imgd.grad
    for name in ("deformNet.conv1.weight", "deformNet.conv3.weight", "deformNet.bn1.weight", "deformNet.bn2.bias", "deformNet.conv4.bias",
                 "fc_reg_ele.weight", "fc_cls_azi.bias", "projector.0.weight", "projector.4.weight", "projector.6.weight",
                 "in/shape_feature", "in/img_feature"):
        want = w_grads[name].numpy()
        assert _rel(got[name].detach().cpu().numpy().reshape(want.shape), want) < 5e-4, name


def _chain_state(Fs, Fi, widths, heads, proj, seed):
    """A PoseEstimator-shaped tail with arbitrary widths (the reference's are 2048/1024/512/200, 24-12-24 x2, 800/400/200)."""
    torch.manual_seed(seed)
    sd = {}
    dims = [Fs + Fi] + list(widths)
    for n in range(1, 5):
        i, o = dims[n - 1], dims[n]
        sd[f"deformNet.conv{n}.weight"] = torch.randn(o, i, 1) / i ** 0.5
        sd[f"deformNet.conv{n}.bias"] = torch.randn(o) * 0.1
        if n < 4:
            sd.update({f"deformNet.bn{n}.weight": torch.randn(o), f"deformNet.bn{n}.bias": torch.randn(o),
                       f"deformNet.bn{n}.running_mean": torch.randn(o) * 0.2, f"deformNet.bn{n}.running_var": torch.rand(o) + 0.5})
    for h, w in zip(pto.HEADS, heads):
        sd[h + ".weight"], sd[h + ".bias"] = torch.randn(w, dims[4]) / dims[4] ** 0.5, torch.randn(w) * 0.1
    pd = [Fi] + list(proj)
    for lin, k in zip((0, 3, 6), range(3)):
        sd[f"projector.{lin}.weight"], sd[f"projector.{lin}.bias"] = torch.randn(pd[k + 1], pd[k]) / pd[k] ** 0.5, torch.randn(pd[k + 1]) * 0.1
    for bn, o in ((1, pd[1]), (4, pd[2])):
        sd.update({f"projector.{bn}.weight": torch.randn(o), f"projector.{bn}.bias": torch.randn(o),
                   f"projector.{bn}.running_mean": torch.randn(o) * 0.2, f"projector.{bn}.running_var": torch.rand(o) + 0.5})
    return sd


@pytest.mark.parametrize("Fs,Fi,B", [(0, 40, 3), (13, 51, 17), (256, 1024, 64), (70, 190, 200)])
def test_ragged_widths_and_batches(pkg, cuda, Fs, Fi, B):
    """Widths that are no multiples of the 64-wide K blocks / 128-wide tiles / 4-wide store groups, an empty shape feature,
    odd head sizes, batches up to the 256-row limit: the zero padding of the operand images and the scalar tails."""
    if Fs == 0:
        pytest.skip("PoseEstimator always concatenates a shape feature")
    sd = _chain_state(Fs, Fi, (Fs + Fi, 97, 66, 50), (7, 5, 9, 7, 5, 9), (131, 70, 33), seed=Fs + B)
    tail = pkg.FrozenPoseTail.from_state_dict(sd).to(cuda)
    g = torch.Generator().manual_seed(B)
    sf, img = torch.randn(B, Fs, generator=g), torch.randn(B, Fi, generator=g)
    for _ in range(2):
        outs, x, p = tail(sf.to(cuda), img.to(cuda))
        w_outs, w_x, w_p = pto.forward(sd, sf, img)
        assert _rel(x.cpu().numpy(), w_x.numpy()) < TOL and _rel(p.cpu().numpy(), w_p.numpy()) < TOL
        assert [tuple(o.shape) for o in outs] == [tuple(o.shape) for o in w_outs]
        assert all(_rel(a.cpu().numpy(), b.numpy()) < TOL for a, b in zip(outs, w_outs))


def test_train_step_frees_its_activations_and_captures_after_eager_steps(pkg, cuda):
    """The autograd node of the train step must not keep itself alive (an output stored as a plain ctx attribute closes a
    node -> tensor -> grad_fn -> node cycle that only a garbage-collector pass breaks): with the collector off, allocated
    memory stays flat over eager steps, and a CUDA-graph capture of the same module right after eager steps on the default
    stream succeeds (with the cycle alive the parameters' AccumulateGrad nodes pin the legacy stream and the capture fails)
    and replays to the same loss and gradients as the eager step."""
    import gc
    torch.manual_seed(5)
    tail = pkg.PoseTail(img_feature_dim=96, shape_feature_dim=32).to(cuda).train()
    B = 24
    sf, img = torch.randn(B, 32, device=cuda), torch.randn(B, 96, device=cuda)

    def fwd_bwd(a, b):
        outs, x, p = tail(a, b)
        loss = sum(o.sum() for o in outs) + x.square().sum() + p.sum()
        loss.backward()
        return loss

    gc.collect()
    gc.disable()
    try:
        a, b = sf.clone().requires_grad_(True), img.clone().requires_grad_(True)
        for _ in range(2):
            tail.zero_grad(set_to_none=True); a.grad = None; b.grad = None
            fwd_bwd(a, b)
        torch.cuda.synchronize()
        m0 = torch.cuda.memory_allocated()
        for _ in range(6):
            tail.zero_grad(set_to_none=True); a.grad = None; b.grad = None
            fwd_bwd(a, b)
        torch.cuda.synchronize()
        assert torch.cuda.memory_allocated() <= m0 + (1 << 16), (torch.cuda.memory_allocated(), m0)
        # reference values of one more eager step from the CURRENT running statistics (train-mode outputs do not depend
        # on them; the gradients neither)
        tail.zero_grad(set_to_none=True); a.grad = None; b.grad = None
        want = fwd_bwd(a, b).item()
        want_g = {n: p_.grad.clone() for n, p_ in tail.named_parameters() if p_.grad is not None}
        want_a = a.grad.clone()
        gs = pkg.GraphedStep(fwd_bwd, (sf.cpu().pin_memory(), img.cpu().pin_memory()), cuda, grad_inputs=(0, 1),
                             zero_grad=lambda: tail.zero_grad(set_to_none=True))
    finally:
        gc.enable()
    gs.stage(sf.cpu().pin_memory(), img.cpu().pin_memory())
    gs.run()
    got = gs.collect()
    assert abs(got - want) <= 1e-5 * abs(want), (got, want)
    assert torch.equal(gs.static[0].grad, want_a)
    for n, p_ in tail.named_parameters():
        if n in want_g:
            assert torch.equal(p_.grad, want_g[n]), n
