"""GPU parity of the point-cloud input producer (package pointcloud.py -> crdpn_pointcloud_sample) against the reference's
read_pointcloud: tests/golden/pointcloud_golden.npz (reference output) and oracle/pointcloud_oracle.py.
Bar: bit-exact for unrotated clouds and for the generated subset indices; rotated clouds within 1 float32 ulp of the span
(the rotation is a float64 matmul on both sides, but cos/sin/FMA contraction may differ in the last float64 bit before the
cast to float32)."""
from pathlib import Path

import numpy as np
import pytest
import torch

from oracle import pointcloud_oracle as pco

pytestmark = pytest.mark.gpu
GOLD = Path(__file__).parent / "golden" / "pointcloud_golden.npz"


@pytest.fixture(scope="module")
def sampler(pkg, cuda):
    return pkg.PointCloudSampler(pco.synthetic_meshes(), point_num=2500, device=cuda, seed=46)


def test_golden_clouds_from_given_subsets(sampler):
    g = np.load(GOLD)
    cases = g["cases"]
    ids = [int(c[0]) for c in cases]
    rots = [float(c[1]) for c in cases]
    subset = np.stack([g[f"subset{j}"] for j in range(len(cases))])
    out = sampler.sample(ids, rots, subset=subset).cpu().numpy()
    for j, rot in enumerate(rots):
        ref = g[f"cloud{j}"]
        if rot == 0:
            assert np.array_equal(out[j], ref), j
        else:
            assert np.abs(out[j] - ref).max() <= 1.2e-7, (j, np.abs(out[j] - ref).max())
            assert (out[j] != ref).mean() < 0.05
        assert out[j].min() == 0.0 and out[j].max() == 1.0


def test_generated_subset_is_the_oracles_permutation(sampler):
    sampler.offset = 5
    ids = [2, 0, 1, 1]
    out, sub = sampler.sample(ids, None, return_subset=True)
    assert sampler.offset == 9
    sub = sub.cpu().numpy()
    meshes = pco.synthetic_meshes()
    for b, cid in enumerate(ids):
        ref_sub = pco.feistel_subset(meshes[cid].shape[0], 2500, 46, 5 + b)
        assert np.array_equal(sub[b], ref_sub), b
        assert len(set(sub[b].tolist())) == 2500
        ref = pco.read_pointcloud(meshes[cid], ref_sub, 0).numpy()
        assert np.array_equal(out[b].cpu().numpy(), ref), b
    # same model twice in one batch: different streams -> different subsets
    assert not np.array_equal(sub[2], sub[3])


@pytest.mark.parametrize("P,V", [(1, 1), (7, 7), (513, 700), (8192, 9000)])
def test_shapes_and_edges(pkg, cuda, P, V):
    mesh = pco.synthetic_meshes((V,), seed=V)[0]
    if P == 1:
        mesh = mesh + 0.0
    s = pkg.PointCloudSampler([mesh], point_num=P, device=cuda, seed=1)
    out, sub = s.sample([0, 0], [0.0, 123.0], return_subset=True)
    sub = sub.cpu().numpy()
    for b, rot in enumerate((0.0, 123.0)):
        assert len(set(sub[b].tolist())) == P and sub[b].max() < V
        ref = pco.read_pointcloud(mesh, sub[b], rot).numpy()
        got = out[b].cpu().numpy()
        if rot == 0:
            assert np.array_equal(got, ref, equal_nan=True)
        else:
            assert np.allclose(got, ref, rtol=0, atol=1.2e-7, equal_nan=True)


def test_errors(pkg, cuda):
    with pytest.raises(ValueError, match="fewer than point_num"):
        pkg.PointCloudSampler([np.zeros((10, 3))], point_num=11, device=cuda)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        pkg.PointCloudSampler([np.zeros((10, 3))], point_num=5, device="cpu")
    s = pkg.PointCloudSampler([np.random.rand(10, 3)], point_num=5, device=cuda)
    with pytest.raises(IndexError):
        s.sample([1])
    with pytest.raises(RuntimeError):
        pkg.PointCloudSampler([np.random.rand(9000, 3)], point_num=8193, device=cuda).sample([0])


def test_feeds_the_encoder(pkg, cuda, sampler):
    """The sampler's output is exactly what ShapeEncoderPC consumes (model.py:257)."""
    enc = pkg.ShapeEncoderPC(256).to(cuda).eval()
    shapes = sampler.sample([0, 1, 2], [0.0, 10.0, 350.0])
    feat = enc(shapes)
    assert feat.shape == (3, 256) and torch.isfinite(feat).all()
