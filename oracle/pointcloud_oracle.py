"""oracle/pointcloud_oracle.py -- CPU restatement of the reference's point-cloud input producer.
TEST INFRASTRUCTURE ONLY (tests/, __graft_entry__.smoke(), bench.py's cpu_baseline leg).

``read_pointcloud`` restates auxiliary/dataset.py:121-150 with the vertex array and the chosen subset as explicit inputs
(the reference reads a mesh file with pymesh and draws the subset from numpy's global RNG).  PINNED:
tests/test_oracle_pointcloud.py runs the reference's own function (pymesh stubbed to return the test's vertices, numpy
seeded so the same subset is drawn) and requires bit equality; golden vectors from that run are in
tests/golden/pointcloud_golden.npz (oracle/gen_golden_pointcloud.py).

``feistel_subset`` restates the kernel's keyed permutation (csrc/pointcloud_sampler.cu) for the case where no subset is given.
"""
from __future__ import annotations

import math

import numpy as np
import torch

M32 = 0xFFFFFFFF


def read_pointcloud(vertices: np.ndarray, subset: np.ndarray, rotation: float = 0) -> torch.Tensor:
    """dataset.py:121-150 for a given subset: [3, P] float32 in [0, 1]."""
    point_cloud = np.asarray(vertices, dtype=np.float64)[np.asarray(subset)]
    if rotation != 0:
        alpha = math.radians(rotation)
        rot_matrix = np.array([[np.cos(alpha), -np.sin(alpha), 0.],
                               [np.sin(alpha), np.cos(alpha), 0.],
                               [0., 0., 1.]])
        point_cloud = np.matmul(point_cloud, rot_matrix.transpose())
    pc = torch.from_numpy(np.ascontiguousarray(point_cloud.transpose())).float()
    pc = pc - torch.min(pc)
    pc = pc / torch.max(pc)
    return pc


def _round(r: int, key: int) -> int:
    x = (r * 0x9E3779B1 + key) & M32
    x ^= x >> 15
    x = (x * 0x85EBCA77) & M32
    x ^= x >> 13
    x = (x * 0xC2B2AE3D) & M32
    x ^= x >> 16
    return x


def feistel_keys(seed: int, stream: int):
    from oracle import crd_oracle
    a = crd_oracle.philox(seed, 2 * stream)
    b = crd_oracle.philox(seed, 2 * stream + 1)
    return [int(a[0]), int(a[1]), int(a[2]), int(a[3]), int(b[0]), int(b[1])]


def feistel_perm(i: int, V: int, keys) -> int:
    bits = 2
    while bits < 64 and (1 << bits) < V:
        bits += 2
    hb = bits // 2
    mask = (1 << hb) - 1
    x = i
    while True:
        L, R = (x >> hb) & mask, x & mask
        for r in range(6):
            L, R = R, L ^ (_round(R, keys[r]) & mask)
        x = (L << hb) | R
        if x < V:
            return x


def feistel_subset(V: int, P: int, seed: int, stream: int) -> np.ndarray:
    keys = feistel_keys(seed, stream)
    return np.array([feistel_perm(i, V, keys) for i in range(P)], dtype=np.int64)


def synthetic_meshes(counts=(3000, 2777, 9001), seed=46):
    """Seeded stand-ins for mesh vertex arrays: float64, anisotropic, off-centre (so min/max normalisation matters)."""
    rng = np.random.default_rng(seed)
    return [rng.normal(size=(v, 3)) * np.array([0.4, 0.15, 0.25]) + np.array([0.1, -0.3, 0.05]) for v in counts]


def call_reference(vertices: np.ndarray, point_num: int, rotation: float, np_seed: int, root: str = "/root/reference"):
    """Run the reference's own read_pointcloud on `vertices` (build container only).  Returns (cloud, subset) or None."""
    import sys
    import types
    from pathlib import Path
    if not Path(root).exists():
        return None
    stubs = {}
    for name in ("matplotlib", "matplotlib.pyplot", "pymesh"):
        if name not in sys.modules:
            sys.modules[name] = types.ModuleType(name)
    sys.modules["matplotlib"].use = lambda *a, **k: None
    sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]
    if root not in sys.path:
        sys.path.insert(0, root)
    try:
        from auxiliary import dataset as ref_ds  # type: ignore
    except Exception:
        return None
    ref_ds.pymesh.load_mesh = lambda path: types.SimpleNamespace(vertices=vertices)
    np.random.seed(np_seed)
    cloud = ref_ds.read_pointcloud("unused.ply", point_num, rotation)
    np.random.seed(np_seed)
    subset = np.random.choice(vertices.shape[0], point_num, replace=False)
    return cloud, subset
