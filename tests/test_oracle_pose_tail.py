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
    g, sd = _gold()
    folded, shape_dim, head_sizes = pkg.FrozenPoseTail.fold_state_dict(sd)
    assert shape_dim == 32 and head_sizes == [24, 12, 24, 24, 12, 24]
    f = {k: v.double() for k, v in folded.items()}
    sf, img = torch.from_numpy(g["in/shape_feature"]).double(), torch.from_numpy(g["in/img_feature"]).double()
    h = torch.relu(f["b1"] + sf @ f["W1t"][:32] + img @ f["W1t"][32:])
    h = torch.relu(f["b2"] + h @ f["W2t"])
    h = torch.relu(f["b3"] + h @ f["W3t"])
    x = torch.tanh(f["b4"] + h @ f["W4t"])
    heads = f["bh"] + x @ f["Wht"]
    p = torch.relu(f["p1"] + img @ f["P1t"])
    p = torch.relu(f["p2"] + p @ f["P2t"])
    p = f["p3"] + p @ f["P3t"]
    outs, ox, op = pto.forward(sd, sf, img)
    assert _rel(x.numpy(), ox.numpy()) < 1e-5 and _rel(p.numpy(), op.numpy()) < 1e-5
    assert _rel(heads.numpy(), torch.cat(outs, 1).numpy()) < 1e-5
