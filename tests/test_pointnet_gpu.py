"""GPU parity tests of the PointNet encoder kernel (eval-mode BN) vs the oracle pinned to the reference's
ShapeEncoderPC (auxiliary/model.py:154-180).  Tolerance (north_star): features within 1e-2 relative in bf16."""
from pathlib import Path

import numpy as np
import pytest
import torch

from oracle import pointnet_oracle as po

pytestmark = pytest.mark.gpu
GOLD = Path(__file__).parent / "golden" / "pointnet_golden.npz"
RELBF = 1e-2


def _rel(got, want):
    return ((got.double() - want.double()).abs().max() / want.double().abs().max()).item()


def _encoder(pkg, st, F, dev):
    enc = pkg.ShapeEncoderPC(F)
    missing = enc.load_state_dict({k: v.clone() for k, v in st.items()})
    assert not missing.missing_keys and not missing.unexpected_keys
    return enc.to(dev).eval()


def test_state_dict_surface_matches_reference(pkg):
    enc = pkg.ShapeEncoderPC(1024)
    sd = enc.state_dict()
    for name, shape in po.PARAM_SHAPES.items():
        assert tuple(sd[name].shape) == tuple(1024 if s == "F" else s for s in shape), name
    assert {k for k in sd if "num_batches_tracked" in k} == {f"bn{i}.num_batches_tracked" for i in (1, 2, 3)}
    assert len(sd) == len(po.PARAM_SHAPES) + 3
    assert sum(p.numel() for p in enc.parameters()) == 143104


def test_golden_vector_from_reference(pkg, cuda):
    g = np.load(GOLD)
    st = {k[6:]: torch.from_numpy(g[k]) for k in g.files if k.startswith("state/")}
    enc = _encoder(pkg, st, 1024, cuda)
    x = torch.from_numpy(g["x"])
    with torch.no_grad():
        out = enc(x.to(cuda)).cpu()
    want = torch.from_numpy(g["eval_out"])            # produced by the reference module itself
    assert out.shape == want.shape and out.dtype == torch.float32
    assert _rel(out, want) < RELBF
    emu = po.forward_bf16_emulated(x, st)             # same precision recipe as the kernel
    assert _rel(out, emu) < 2e-3


@pytest.mark.parametrize("B,P,F", [(1, 1, 1024), (2, 7, 128), (3, 128, 256), (2, 256, 512), (5, 257, 1024),
                                    (4, 1000, 1024), (149, 300, 1024), (2, 2500, 1024), (1, 5000, 1024)])
def test_eval_parity_shapes(pkg, cuda, B, P, F):
    st = po.random_state(F, seed=B * 7 + P)
    x = po.random_clouds(B, P, seed=P)
    enc = _encoder(pkg, st, F, cuda)
    out = enc(x.to(cuda)).cpu()
    want = po.forward(x, st, training=False)
    assert out.shape == (B, F)
    assert _rel(out, want) < RELBF
    assert _rel(out, po.forward_bf16_emulated(x, st)) < 2e-3


def test_permutation_and_duplication_invariance_bitwise(pkg, cuda):
    st = po.random_state(1024, seed=1)
    x = po.random_clouds(3, 700, seed=2).to(cuda)
    enc = _encoder(pkg, st, 1024, cuda)
    a = enc(x)
    perm = torch.randperm(700, device=cuda)
    b = enc(x[:, :, perm].contiguous())               # max over points: order of points is irrelevant
    c = enc(torch.cat([x, x], dim=2))                 # idempotence: every point twice
    assert torch.equal(a, b) and torch.equal(a, c)
    d = enc(x[1:2].contiguous())                      # clouds are independent
    assert torch.equal(a[1:2], d)


def test_repack_after_parameter_change_and_cpu_input_is_loud(pkg, cuda):
    st = po.random_state(1024, seed=3)
    x = po.random_clouds(2, 300, seed=4)
    enc = _encoder(pkg, st, 1024, cuda)
    a = enc(x.to(cuda))
    with torch.no_grad():
        enc.bn3.bias.add_(1.0)
    b = enc(x.to(cuda))
    assert torch.allclose(b, a + 1.0, atol=1e-5)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        enc.eval()(x)


def test_config2_full_size(pkg, cuda):
    """BASELINE configs[1]: batch 160, 2500 points, 3->64->128->1024 + max-pool."""
    st = po.random_state(1024, seed=46)
    x = po.random_clouds(160, 2500, seed=46)
    enc = _encoder(pkg, st, 1024, cuda)
    out = enc(x.to(cuda)).cpu()
    want = po.forward(x, st, training=False, dtype=torch.float32)
    assert _rel(out, want) < RELBF
