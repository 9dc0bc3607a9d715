"""CPU tests pinning oracle/pose_tail_oracle.py to the reference's PoseEstimator tail (auxiliary/model.py:183-203, 238-272):
against tests/golden/pose_tail_golden.npz (made from the reference's own modules) and, where /root/reference is mounted,
against those modules live.  Also: the BatchNorm folding / head concatenation the GPU module uses is exact algebra (checked
here on CPU tensors through the module's own fold function; its forward needs a GPU)."""
from pathlib import Path

import numpy as np
import pytest
import torch

from oracle import pose_tail_oracle as pto

GOLD = Path(__file__).parent / "golden" / "pose_tail_golden.npz"


def _gold():
    g = np.load(GOLD)
    sd = {k[6:]: torch.from_numpy(g[k]) for k in g.files if k.startswith("state/")}
    return g, sd


def _rel(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return np.abs(a - b).max() / (np.abs(b).max() + 1e-30)


def test_oracle_matches_golden():
    g, sd = _gold()
    outs, x, p = pto.forward(sd, torch.from_numpy(g["in/shape_feature"]), torch.from_numpy(g["in/img_feature"]))
    assert _rel(x.numpy(), g["out/x"]) < 1e-5 and _rel(p.numpy(), g["out/projector"]) < 1e-5
    for i, o in enumerate(outs):
        assert _rel(o.numpy(), g[f"out/head{i}"]) < 1e-5
    assert [o.shape[1] for o in outs] == [24, 12, 24, 24, 12, 24]


@pytest.mark.skipif(not Path("/root/reference/auxiliary/model.py").exists(), reason="reference not mounted")
def test_oracle_matches_live_reference_modules():
    m, sd = pto.reference_tail(img_dim=48, shape_dim=16, seed=3)
    g = torch.Generator().manual_seed(1)
    sf, img = torch.randn(5, 16, generator=g), torch.randn(5, 48, generator=g)
    r_outs, r_x, r_p = pto.reference_forward(m, sf, img)
    outs, x, p = pto.forward(sd, sf, img)
    assert _rel(x.numpy(), r_x.numpy()) < 1e-5 and _rel(p.numpy(), r_p.numpy()) < 1e-5
    assert all(_rel(a.numpy(), b.numpy()) < 1e-5 for a, b in zip(outs, r_outs))


def test_folding_is_exact_algebra(pkg):
    """The folded fp32 weights / biases the GPU module packs (FrozenPoseTail buffers W<l>, b<l>, s<l>: BatchNorm as a row
    scale, the six heads as one layer, the concat as the first layer's input range) reproduce the oracle on CPU tensors."""
    g, sd = _gold()
    tail = pkg.FrozenPoseTail.from_state_dict(sd)
    assert tail.shape_dim == 32 and tail.img_dim == 64 and tail.head_sizes == [24, 12, 24, 24, 12, 24]
    sf, img = torch.from_numpy(g["in/shape_feature"]).double(), torch.from_numpy(g["in/img_feature"]).double()
    vals = {-1: torch.cat((sf, img), 1), -2: img}
    for l, s in enumerate(tail.spec):
        W = getattr(tail, f"W{l}").double() * getattr(tail, f"s{l}").double()[:, None]
        z = vals[s["src"]] @ W.t() + getattr(tail, f"b{l}").double()
        vals[l] = torch.relu(z) if s["act"] == 1 else torch.tanh(z) if s["act"] == 2 else z
    outs, ox, op = pto.forward(sd, sf, img)
    assert _rel(vals[3].numpy(), ox.numpy()) < 1e-5 and _rel(vals[7].numpy(), op.numpy()) < 1e-5
    assert _rel(vals[4].numpy(), torch.cat(outs, 1).numpy()) < 1e-5


def _train_gold(g):
    g_outs = [torch.from_numpy(g[f"train/gin/head{i}"]) for i in range(6)]
    return g_outs, torch.from_numpy(g["train/gin/x"]), torch.from_numpy(g["train/gin/projector"])


def test_train_oracle_matches_golden():
    """Train mode (training.py:30,47,75): outputs, updated running statistics and every gradient of the fp64 oracle against
    what the reference's own modules produced (fp32) for the same inputs and upstream gradients."""
    g, sd = _gold()
    g_outs, g_x, g_p = _train_gold(g)
    outs, x, p, new_running, grads = pto.train_step_with_grads(sd, torch.from_numpy(g["in/shape_feature"]),
                                                               torch.from_numpy(g["in/img_feature"]), g_outs, g_x, g_p)
    assert _rel(x.numpy(), g["train/out/x"]) < 1e-5 and _rel(p.numpy(), g["train/out/projector"]) < 1e-5
    for i, o in enumerate(outs):
        assert _rel(o.numpy(), g[f"train/out/head{i}"]) < 1e-5
    for k, v in new_running.items():
        assert _rel(v.numpy(), g["train/state/" + k]) < 1e-5, k
    checked = 0
    for k in g.files:
        if k.startswith("train/grad/"):
            name = k[len("train/grad/"):]
            want = g[k]
            got = grads[name].numpy().reshape(want.shape)
            scale = np.abs(want).max()
            if ".conv" in name and name.endswith(".bias") and "conv4" not in name or name in ("projector.0.bias", "projector.3.bias"):
                assert np.abs(got).max() < 1e-9 and scale < 1e-4   # a bias in front of train-mode BatchNorm has zero gradient
            else:
                assert _rel(got, want) < 2e-4, name
            checked += 1
    assert checked == 38   # 7 Linear x 2 + 5 BatchNorm x 2 + 6 heads x 2 + the two inputs


@pytest.mark.skipif(not Path("/root/reference/auxiliary/model.py").exists(), reason="reference not mounted")
def test_train_oracle_matches_live_reference_modules():
    m, sd = pto.reference_tail(img_dim=48, shape_dim=16, seed=3)
    gen = torch.Generator().manual_seed(2)
    sf, img = torch.randn(7, 16, generator=gen), torch.randn(7, 48, generator=gen)
    g_outs = [torch.randn(7, w, generator=gen) for w in (24, 12, 24, 24, 12, 24)]
    g_x, g_p = torch.randn(7, 200, generator=gen), torch.randn(7, 200, generator=gen)
    outs, x, p, new_running, grads = pto.train_step_with_grads(sd, sf, img, g_outs, g_x, g_p)
    r_outs, r_x, r_p, r_grads = pto.reference_train_step(m, sf, img, g_outs, g_x, g_p)
    assert _rel(x.numpy(), r_x.numpy()) < 1e-5 and _rel(p.numpy(), r_p.numpy()) < 1e-5
    for name in ("deformNet.conv1.weight", "deformNet.bn2.weight", "fc_cls_azi.weight", "projector.3.weight", "in/shape_feature",
                 "in/img_feature"):
        assert _rel(grads[name].numpy().reshape(r_grads[name].shape), r_grads[name].numpy()) < 2e-4, name
    msd = m.state_dict()
    for k, v in new_running.items():
        assert _rel(v.numpy(), msd[k].numpy()) < 1e-5, k
