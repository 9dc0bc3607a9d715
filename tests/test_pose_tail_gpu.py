"""GPU parity of FrozenPoseTail (folded weights, split-K concat, concatenated heads, one CUDA graph) against the reference's
own outputs (tests/golden/pose_tail_golden.npz) and the oracle: fp32, <= 1e-5 of each tensor's max."""
from pathlib import Path

import numpy as np
import pytest
import torch

from oracle import pose_tail_oracle as pto

pytestmark = pytest.mark.gpu
GOLD = Path(__file__).parent / "golden" / "pose_tail_golden.npz"


def _rel(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return np.abs(a - b).max() / (np.abs(b).max() + 1e-30)


@pytest.mark.parametrize("graph", [False, True])
def test_matches_reference_golden(pkg, cuda, graph):
    g = np.load(GOLD)
    sd = {k[6:]: torch.from_numpy(g[k]) for k in g.files if k.startswith("state/")}
    tail = pkg.FrozenPoseTail.from_state_dict(sd, graph=graph).to(cuda)
    sf, img = torch.from_numpy(g["in/shape_feature"]).to(cuda), torch.from_numpy(g["in/img_feature"]).to(cuda)
    for _ in range(2):      # second call replays the captured graph
        outs, x, p = tail(sf, img)
        assert _rel(x.cpu().numpy(), g["out/x"]) < 1e-5 and _rel(p.cpu().numpy(), g["out/projector"]) < 1e-5
        for i, o in enumerate(outs):
            assert o.shape == g[f"out/head{i}"].shape and _rel(o.cpu().numpy(), g[f"out/head{i}"]) < 1e-5


def test_reference_size_and_new_inputs_through_the_graph(pkg, cuda):
    """The KD-time shapes (138 rows, 1024 + 1024 features) with random weights vs the oracle; the captured graph must pick
    up NEW inputs on replay, and a different batch size captures its own graph."""
    torch.manual_seed(0)
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
    tail = pkg.FrozenPoseTail.from_state_dict(sd).to(cuda)
    for B, seed in ((138, 1), (138, 2), (46, 3)):
        g = torch.Generator().manual_seed(seed)
        sf, img = torch.randn(B, 1024, generator=g), torch.randn(B, 1024, generator=g)
        outs, x, p = tail(sf.to(cuda), img.to(cuda))
        w_outs, w_x, w_p = pto.forward(sd, sf, img)
        assert _rel(x.cpu().numpy(), w_x.numpy()) < 2e-5 and _rel(p.cpu().numpy(), w_p.numpy()) < 2e-5
        assert all(_rel(a.cpu().numpy(), b.numpy()) < 2e-5 for a, b in zip(outs, w_outs))
    assert len(tail._graphs) == 2
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        tail(torch.zeros(2, 1024), torch.zeros(2, 1024))
    # bf16 mode: north_star's 1e-2 tolerance
    tail16 = pkg.FrozenPoseTail.from_state_dict(sd, dtype=torch.bfloat16).to(cuda)
    g = torch.Generator().manual_seed(9)
    sf, img = torch.randn(138, 1024, generator=g), torch.randn(138, 1024, generator=g)
    outs, x, p = tail16(sf.to(cuda), img.to(cuda))
    w_outs, w_x, w_p = pto.forward(sd, sf, img)
    assert x.dtype == torch.float32
    assert _rel(x.cpu().numpy(), w_x.numpy()) < 2e-2 and _rel(p.cpu().numpy(), w_p.numpy()) < 2e-2
    assert all(_rel(a.cpu().numpy(), b.numpy()) < 2e-2 for a, b in zip(outs, w_outs))
