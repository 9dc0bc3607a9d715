"""Point-cloud input producer -- batch replacement for the reference's ``read_pointcloud`` on a B200 kernel.

Reference: ``auxiliary/dataset.py:121-150`` (called once per sample inside DataLoader workers, ``dataset.py:299,607``):
``pymesh.load_mesh(path).vertices`` -> ``np.random.choice(V, point_num, replace=False)`` -> optional rotation about z ->
``[3, P]`` float32 -> minus global min, divided by global max.

``PointCloudSampler`` keeps the raw vertices of every model resident in HBM (float64, as pymesh yields them) and produces
a step's ``[B, 3, P]`` batch in one launch (``crdpn_pointcloud_sample``), ready for ``ShapeEncoderPC``.  The subset is
either passed in (``subset=``; then the result equals ``read_pointcloud`` for that subset bit for bit when unrotated) or
drawn in the kernel from a keyed Feistel permutation (distinct by construction; reproducible from ``seed``).
Mesh parsing stays on the host and out of scope: construct the sampler from arrays.
"""
from __future__ import annotations

import torch

from . import _native


class PointCloudSampler:
    def __init__(self, vertex_arrays, point_num: int = 2500, device="cuda", seed: int | None = None):
        """vertex_arrays: sequence of [V_m, 3] arrays / tensors (model m's mesh vertices)."""
        device = torch.device(device)
        if device.type != "cuda":
            raise RuntimeError("PointCloudSampler needs a CUDA device: this package has no CPU fallback")
        vs = [torch.as_tensor(v, dtype=torch.float64).reshape(-1, 3) for v in vertex_arrays]
        if not vs:
            raise ValueError("no point clouds given")
        counts = [int(v.shape[0]) for v in vs]
        if min(counts) < point_num:
            raise ValueError(f"a model has {min(counts)} vertices, fewer than point_num={point_num} "
                             "(the reference's np.random.choice(..., replace=False) raises here too)")
        self.point_num = int(point_num)
        self.counts = counts
        self.vertices = torch.cat(vs, 0).contiguous().to(device)
        self.offsets = torch.tensor([0] + list(torch.tensor(counts).cumsum(0).tolist()), dtype=torch.int64, device=device)
        self.seed = int(torch.initial_seed() if seed is None else seed) & 0xFFFFFFFFFFFFFFFF
        self.offset = 0
        self.device = device

    def __len__(self):
        return len(self.counts)

    def sample(self, cloud_ids, rotations=None, subset=None, return_subset=False):
        """cloud_ids [B] (model index per sample), rotations [B] degrees or None, subset [B, P] int64 or None.
        Returns shapes [B, 3, P] float32 (and the chosen vertex rows [B, P] when return_subset)."""
        dev = self.device
        ids = torch.as_tensor(cloud_ids, dtype=torch.int64)
        if not ids.is_cuda and ids.numel() and (int(ids.min()) < 0 or int(ids.max()) >= len(self.counts)):
            raise IndexError("cloud id out of range")   # (device-resident ids are not read back: no hidden sync)
        ids = ids.to(dev).contiguous()
        B, P = ids.numel(), self.point_num
        rot = None
        if rotations is not None:
            rot = torch.as_tensor(rotations, dtype=torch.float32).to(dev).contiguous()
            if rot.numel() != B:
                raise RuntimeError("rotations must have one entry per cloud")
        sub = None
        if subset is not None:
            sub = torch.as_tensor(subset, dtype=torch.int64).to(dev).contiguous()
            if sub.shape != (B, P):
                raise RuntimeError(f"subset must be [B, point_num] = {(B, P)}")
        out = torch.empty(B, 3, P, dtype=torch.float32, device=dev)
        sub_out = torch.empty(B, P, dtype=torch.int64, device=dev) if return_subset else None
        with _native.on_device(dev):
            rc = _native.lib().crdpn_pointcloud_sample(
                self.vertices.data_ptr(), self.offsets.data_ptr(), ids.data_ptr(), rot.data_ptr() if rot is not None else None,
                sub.data_ptr() if sub is not None else None, self.seed, self.offset, B, P, out.data_ptr(),
                sub_out.data_ptr() if sub_out is not None else None, _native.stream_ptr(dev))
        _native.check(rc, "crdpn_pointcloud_sample")
        if sub is None:
            self.offset += B
        return (out, sub_out) if return_subset else out
