"""CPU tests pinning oracle/pointcloud_oracle.py to the reference's read_pointcloud (auxiliary/dataset.py:121-150):
bit equality against tests/golden/pointcloud_golden.npz (made from the reference by oracle/gen_golden_pointcloud.py) and,
where /root/reference is mounted, against the reference function itself; plus the keyed-permutation subset's properties."""
from pathlib import Path

import numpy as np
import pytest

from oracle import pointcloud_oracle as pco

GOLD = Path(__file__).parent / "golden" / "pointcloud_golden.npz"


def test_read_pointcloud_matches_golden_bit_for_bit():
    g = np.load(GOLD)
    meshes = pco.synthetic_meshes()
    assert [m.shape[0] for m in meshes] == g["counts"].tolist()
    for j, (cid, rot, _) in enumerate(g["cases"]):
        out = pco.read_pointcloud(meshes[int(cid)], g[f"subset{j}"], float(rot)).numpy()
        assert np.array_equal(out, g[f"cloud{j}"]), j
        assert out.min() == 0.0 and out.max() == 1.0 and out.shape == (3, 2500)


@pytest.mark.skipif(not Path("/root/reference/auxiliary/dataset.py").exists(), reason="reference not mounted")
@pytest.mark.parametrize("rot", [0, 45.5, 359])
def test_against_live_reference(rot):
    mesh = pco.synthetic_meshes((4000,), seed=3)[0]
    r = pco.call_reference(mesh, 1000, rot, 21)
    assert r is not None
    cloud, subset = r
    assert len(set(subset.tolist())) == 1000
    assert np.array_equal(pco.read_pointcloud(mesh, subset, rot).numpy(), cloud.numpy())


@pytest.mark.parametrize("V,P", [(2500, 2500), (2501, 2500), (3000, 2500), (65536, 2500), (65537, 100), (1, 1), (5, 5)])
def test_feistel_subset_is_distinct_and_in_range(V, P):
    s = pco.feistel_subset(V, P, seed=46, stream=V)
    assert s.shape == (P,) and s.min() >= 0 and s.max() < V and len(set(s.tolist())) == P
    if V == P:
        assert sorted(s.tolist()) == list(range(V))   # a permutation of the whole cloud


def test_feistel_subset_is_keyed_and_roughly_uniform():
    a = pco.feistel_subset(3000, 2500, 46, 0)
    b = pco.feistel_subset(3000, 2500, 46, 1)
    c = pco.feistel_subset(3000, 2500, 47, 0)
    assert not np.array_equal(a, b) and not np.array_equal(a, c)
    assert np.array_equal(a, pco.feistel_subset(3000, 2500, 46, 0))
    # inclusion frequency of every vertex over 200 streams: mean P/V = 1/6 at V = 600, P = 100
    V, P, runs = 600, 100, 200
    hits = np.zeros(V)
    first = np.zeros(V)
    for s in range(runs):
        sub = pco.feistel_subset(V, P, 9, s)
        hits[sub] += 1
        first[sub[0]] += 1
    exp = runs * P / V
    chi2 = ((hits - exp) ** 2 / (exp * (1 - P / V))).sum()   # ~ chi-square with V-1 dof: mean 599, sd ~35
    assert 450 < chi2 < 760, chi2
    assert first.max() <= 6   # no vertex is a favourite first pick
