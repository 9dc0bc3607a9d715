"""PointNet point-cloud encoder -- drop-in for the reference's ``ShapeEncoderPC`` on B200 kernels.

Mirrors ``auxiliary/model.py:154-180``: same constructor (``feature_dim=1024``), same sub-module /
parameter / buffer names and shapes (``conv1.weight[64,3,1]`` ... ``bn3.running_var[F]``,
``num_batches_tracked``), same default initialisation (the holders ARE ``nn.Conv1d`` / ``nn.BatchNorm1d``),
so checkpoints load by key under the ``shape_encoder.`` prefix (``auxiliary/utils.py:56-73``) and the module
slots into ``PoseEstimator`` at ``model.py:222-223,257``.

forward(shapes[B,3,P] float32) -> [B, feature_dim] float32.

* ``.eval()`` (the KD-time teacher, ``KD/common/base_class.py:317``): one fused sm_100a kernel
  (``crdpn_pointnet_forward_eval``): BN folded into the weights, layers 2-3 on tcgen05 tensor cores in bf16
  with fp32 accumulation, max over points fused into the GEMM epilogue.  The output carries no autograd
  graph (the teacher is frozen in the KD loop; its gradients are never consumed).
* ``.train()`` (teacher training, ``training.py:30,47,75``): batch-statistics BatchNorm forward
  (``crdpn_pointnet_forward_train``: analytic BN1 statistics, a tensor-core statistics pass for BN2, the fused
  kernel with BN3's statistics / max / arg-max in the epilogue) and its backward (``crdpn_pointnet_backward``),
  wired into autograd for the 12 parameters; running statistics are updated in place like ``nn.BatchNorm1d``.
  ``train_precision = "fp32"`` (default) evaluates every tensor-core product as three fp16 hi/lo MMAs so that the
  arg-max points and ReLU gates that route the gradients are those of the fp32 reference; ``"bf16"`` is the
  single-MMA recipe (features within 1e-2, gradients re-routed at near-ties).

No CPU fallback: inputs must be CUDA tensors.
"""
from __future__ import annotations

import ctypes

import torch
from torch import nn

from . import _native

BN_EPS = 1e-5


class ShapeEncoderPC(nn.Module):
    """Shape Encoder using point cloud (reference docstring: returns a tensor of size NxC)."""

    def __init__(self, feature_dim: int = 1024):
        super().__init__()
        self.conv1 = torch.nn.Conv1d(3, 64, 1)
        self.conv2 = torch.nn.Conv1d(64, 128, 1)
        self.conv3 = torch.nn.Conv1d(128, feature_dim, 1)
        self.bn1 = torch.nn.BatchNorm1d(64)
        self.bn2 = torch.nn.BatchNorm1d(128)
        self.bn3 = torch.nn.BatchNorm1d(feature_dim)
        self.feature_dim = feature_dim
        self.variant = 0
        # train-mode arithmetic: "fp32" = every tensor-core product as three fp16 hi/lo MMAs (conv outputs within ~2e-7
        # of an fp32 run, so arg-max points and ReLU gates -- hence the parameter gradients -- are the reference's);
        # "bf16" = one bf16 MMA per product (features within 1e-2, faster, gradients re-routed at near-ties)
        self.train_precision = "fp32"
        self._sync = None   # (process group, world size) once sync_batchnorm() is called
        self._sync_comm, self._sync_equal, self._sync_px = "dist", False, None
        self._packed = None
        self._packed_key = None
        self._ws = None

    # -- eval path -------------------------------------------------------------------------------------
    def _pack_sources(self):
        return [self.conv1.weight, self.conv1.bias, self.conv2.weight, self.conv2.bias, self.conv3.weight,
                self.conv3.bias,
                self.bn1.weight, self.bn1.bias, self.bn1.running_mean, self.bn1.running_var,
                self.bn2.weight, self.bn2.bias, self.bn2.running_mean, self.bn2.running_var,
                self.bn3.weight, self.bn3.bias, self.bn3.running_mean, self.bn3.running_var]

    def _packed_params(self, device):
        """bf16 tensor-core operand images with eval-mode BN folded in; rebuilt whenever a source tensor changes."""
        src = self._pack_sources()
        key = (device,) + tuple((t.data_ptr(), t._version) for t in src)
        if key != self._packed_key:
            for t in src:
                if t.device != device or t.dtype != torch.float32:
                    raise RuntimeError("ShapeEncoderPC parameters must be float32 on the input's CUDA device")
            n = ctypes.c_size_t(0)
            _native.check(_native.lib().crdpn_pointnet_packed_bytes(self.feature_dim, ctypes.byref(n)),
                          "crdpn_pointnet_packed_bytes")
            if self._packed is None or self._packed.numel() != n.value or self._packed.device != device:
                self._packed = torch.empty(n.value, dtype=torch.uint8, device=device)
            keep = [t.detach().contiguous() for t in src]   # alive until the pack launch has been enqueued
            with _native.on_device(device):
                rc = _native.lib().crdpn_pointnet_pack(*[t.data_ptr() for t in keep], BN_EPS, self.feature_dim,
                                                       self._packed.data_ptr(), _native.stream_ptr(device))
            _native.check(rc, "crdpn_pointnet_pack")
            del keep
            self._packed_key = key
        return self._packed

    def forward_eval(self, shapes: torch.Tensor) -> torch.Tensor:
        if not shapes.is_cuda:
            raise RuntimeError("ShapeEncoderPC input must be a CUDA tensor: this package has no CPU fallback")
        if shapes.dim() != 3 or shapes.shape[1] != 3:
            raise RuntimeError(f"expected shapes[B,3,P], got {tuple(shapes.shape)}")
        x = shapes.detach().to(torch.float32).contiguous()
        B, _, P = x.shape
        dev = x.device
        packed = self._packed_params(dev)
        n = ctypes.c_size_t(0)
        _native.check(_native.lib().crdpn_pointnet_workspace_bytes(B, P, self.feature_dim, dev.index or 0,
                                                                   ctypes.byref(n)), "crdpn_pointnet_workspace_bytes")
        if self._ws is None or self._ws.numel() < n.value or self._ws.device != dev:
            self._ws = torch.empty(n.value, dtype=torch.uint8, device=dev)
        out = torch.empty(B, self.feature_dim, dtype=torch.float32, device=dev)
        with _native.on_device(dev):
            rc = _native.lib().crdpn_pointnet_forward_eval(x.data_ptr(), B, P, self.feature_dim, packed.data_ptr(),
                                                           out.data_ptr(), self._ws.data_ptr(), self._ws.numel(),
                                                           self.variant, _native.stream_ptr(dev))
        _native.check(rc, "crdpn_pointnet_forward_eval")
        return out.view(-1, self.feature_dim)

    def forward(self, shapes: torch.Tensor) -> torch.Tensor:
        if self.training:
            return self.forward_train(shapes)
        return self.forward_eval(shapes)

    # -- train path ------------------------------------------------------------------------------------
    _PARAM_ORDER = ("conv1.weight", "conv1.bias", "conv2.weight", "conv2.bias", "conv3.weight", "conv3.bias",
                    "bn1.weight", "bn1.bias", "bn2.weight", "bn2.bias", "bn3.weight", "bn3.bias")

    def _train_params(self):
        return [self.conv1.weight, self.conv1.bias, self.conv2.weight, self.conv2.bias, self.conv3.weight,
                self.conv3.bias, self.bn1.weight, self.bn1.bias, self.bn2.weight, self.bn2.bias,
                self.bn3.weight, self.bn3.bias]

    def sync_batchnorm(self, group=None, enabled: bool = True, comm: str = "dist", equal_batches: bool = False):
        """Train-mode statistics over the clouds of ALL ranks of `group` (SyncBN-style).

        Forward and backward then run in phases with a sum over ranks of the per-channel accumulators in between
        (``crdpn_pointnet_sync_blocks``: <= 2F doubles per hand-off, three hand-offs each way).  The parameter gradients
        returned on every rank are the gradients of the SUM over ranks of the local losses and are identical on all
        ranks -- do not average them again (no DDP wrapper); scale the loss by 1/world_size for a global mean.

        comm="dist": the hand-offs are torch.distributed all-reduces on the kernels' own memory (10 small collectives per
        step).  comm="p2p": each hand-off is ONE kernel over NVLink peer memory (``crdpn_p2p_allreduce_blocks``: every
        element is pushed into each peer's slot and summed in rank order, in place; 6 launches per step, no NCCL call; the
        ranks must stay within CRDPN_P2P_TIMEOUT_S of each other).
        equal_batches=True promises that every rank feeds the same number of points per step, so the global point count
        is world_size x local instead of one more all-reduce and host read per forward."""
        import torch.distributed as dist
        if comm not in ("dist", "p2p"):
            raise ValueError("comm must be 'dist' or 'p2p'")
        self._sync = (group, dist.get_world_size(group)) if enabled else None
        self._sync_comm = comm
        self._sync_equal = bool(equal_batches)
        self._sync_px = None
        return self

    def _sync_exchange(self, device):
        """Peer-memory exchange buffers of the hand-offs (collective set-up on first use)."""
        if self._sync_px is None:
            import torch.distributed as dist
            from .sharded import PeerExchange
            group = self._sync[0]
            # a reduction slot holds 2*Bmax*Dmax + 8 + 8*Bmax 32-bit words (a double takes two).  Largest hand-offs
            # (crdpn_pointnet_sync_blocks): backward phase 0 = 128 doubles + F*128 + 2F floats; backward phase 1 = 256
            # doubles + 128*64 + 2*128*128 floats
            words = max(self.feature_dim * 128 + 2 * self.feature_dim + 256, 128 * 64 + 2 * 128 * 128 + 512)
            bmax = max(64, (words + 255) // 256 + 1)
            self._sync_px = PeerExchange(group, dist.get_rank(group), self._sync[1], device, bmax, 128)
        return self._sync_px

    def forward_train(self, shapes: torch.Tensor) -> torch.Tensor:
        """Batch-statistics BatchNorm forward (``training.py:30,47``); differentiable w.r.t. the 12 parameters
        (``training.py:75``); running statistics and ``num_batches_tracked`` are updated in place."""
        if not shapes.is_cuda:
            raise RuntimeError("ShapeEncoderPC input must be a CUDA tensor: this package has no CPU fallback")
        if shapes.dim() != 3 or shapes.shape[1] != 3:
            raise RuntimeError(f"expected shapes[B,3,P], got {tuple(shapes.shape)}")
        if shapes.requires_grad:
            raise RuntimeError("ShapeEncoderPC: gradients w.r.t. the input point cloud are not provided (it is data)")
        for bn in (self.bn1, self.bn2, self.bn3):
            if bn.momentum is None or not bn.track_running_stats or not bn.affine:
                raise RuntimeError("ShapeEncoderPC train mode expects nn.BatchNorm1d defaults (momentum, running stats, affine)")
        if self.train_precision not in ("fp32", "bf16"):
            raise RuntimeError("ShapeEncoderPC.train_precision must be 'fp32' or 'bf16'")
        return _PointNetTrainFunction.apply(self, shapes, *self._train_params()).view(-1, self.feature_dim)


def _aligned(nbytes: int, device) -> tuple[torch.Tensor, int]:
    """A uint8 buffer with a 1024-byte aligned start; returns (owner tensor, aligned pointer)."""
    buf = torch.empty(nbytes + 1024, dtype=torch.uint8, device=device)
    return buf, (buf.data_ptr() + 1023) & ~1023


def _global_points(local_points: int, device, group) -> int:
    import torch.distributed as dist
    t = torch.tensor([local_points], dtype=torch.int64, device=device)
    dist.all_reduce(t, group=group)
    return int(t.item())


def _sum_over_ranks(lib, dims, sync_point: int, group, buffers, px=None) -> None:
    """All-reduce (SUM) the accumulators named by crdpn_pointnet_sync_blocks(sync_point).  `buffers`: id -> (owner
    tensor, base pointer); the blocks are viewed in place through the owner's storage, so the collective runs on the
    kernels' own memory on the current stream.  With a PeerExchange `px` all blocks of the hand-off are summed by one
    kernel over NVLink peer memory instead."""
    import torch.distributed as dist
    nb = ctypes.c_int(0)
    buf, off, cnt, f64 = (ctypes.c_int * 4)(), (ctypes.c_size_t * 4)(), (ctypes.c_int64 * 4)(), (ctypes.c_int * 4)()
    _native.check(lib.crdpn_pointnet_sync_blocks(*dims, sync_point, ctypes.byref(nb), buf, off, cnt, f64), "crdpn_pointnet_sync_blocks")
    if px is not None:
        ptrs = (ctypes.c_void_p * 4)(*[buffers[buf[i]][1] + off[i] if i < nb.value else None for i in range(4)])
        dev = px.device
        with _native.on_device(dev):
            rc = lib.crdpn_p2p_allreduce_blocks(ptrs, cnt, f64, nb.value, px._ptrs, px.rank, px.world, px.Bmax, px.Dmax,
                                                _native.stream_ptr(dev))
        _native.check(rc, "crdpn_p2p_allreduce_blocks")
        return
    for i in range(nb.value):
        owner, base = buffers[buf[i]]
        flat = owner.view(-1)
        start = base - flat.data_ptr() + off[i]                      # byte offset inside the owner tensor
        esz = flat.element_size()
        nbytes = cnt[i] * (8 if f64[i] else 4)
        view = flat[start // esz:(start + nbytes) // esz].view(torch.float64 if f64[i] else torch.float32)
        dist.all_reduce(view, group=group)


class _PointNetTrainFunction(torch.autograd.Function):
    """ctypes glue around ``crdpn_pointnet_forward_train`` / ``crdpn_pointnet_backward``."""

    @staticmethod
    def forward(ctx, module, shapes, *params):
        lib = _native.lib()
        x = shapes.detach().to(torch.float32).contiguous()
        B, _, P = x.shape
        F = module.feature_dim
        dev = x.device
        for t in params:
            if t.device != dev or t.dtype != torch.float32:
                raise RuntimeError("ShapeEncoderPC parameters must be float32 on the input's CUDA device")
        p = [t.detach().contiguous() for t in params]
        n = ctypes.c_size_t(0)
        _native.check(lib.crdpn_pointnet_train_ctx_bytes(B, P, F, ctypes.byref(n)), "crdpn_pointnet_train_ctx_bytes")
        owner, cptr = _aligned(n.value, dev)
        out = torch.empty(B, F, dtype=torch.float32, device=dev)
        bns = (module.bn1, module.bn2, module.bn3)
        args = [x.data_ptr(), B, P, F] + [t.data_ptr() for t in p[:6]]
        for i, bn in enumerate(bns):
            args += [p[6 + 2 * i].data_ptr(), p[7 + 2 * i].data_ptr(), bn.running_mean.data_ptr(),
                     bn.running_var.data_ptr(), bn.num_batches_tracked.data_ptr()]
        sync = module._sync
        variant = (module.variant & ~16) | (16 if module.train_precision == "bf16" else 0)
        total = B * P
        px = None
        if sync is not None:
            total = B * P * sync[1] if module._sync_equal else _global_points(B * P, dev, sync[0])
            px = module._sync_exchange(dev) if module._sync_comm == "p2p" else None
        with _native.on_device(dev):
            if sync is None:
                rc = lib.crdpn_pointnet_forward_train(*args, float(module.bn1.eps), float(module.bn1.momentum),
                                                      out.data_ptr(), cptr, n.value, variant,
                                                      _native.stream_ptr(dev))
                _native.check(rc, "crdpn_pointnet_forward_train")
            else:
                for ph in range(4):
                    rc = lib.crdpn_pointnet_forward_train_phased(*args, float(module.bn1.eps), float(module.bn1.momentum),
                                                                 out.data_ptr(), cptr, n.value, variant, ph, ph + 1,
                                                                 total, _native.stream_ptr(dev))
                    _native.check(rc, "crdpn_pointnet_forward_train_phased")
                    if ph < 3:
                        _sum_over_ranks(lib, (B, P, F), ph, sync[0], {0: (owner, cptr)}, px)
        ctx.save_for_backward(x, *p)
        ctx.train_ctx = (owner, cptr, n.value)
        ctx.dims = (B, P, F)
        ctx.sync = (sync, total, px)
        return out

    @staticmethod
    def backward(ctx, grad_out):
        lib = _native.lib()
        x, *p = ctx.saved_tensors
        owner, cptr, cbytes = ctx.train_ctx
        B, P, F = ctx.dims
        dev = x.device
        g = grad_out.detach().to(torch.float32).contiguous()
        n = ctypes.c_size_t(0)
        _native.check(lib.crdpn_pointnet_backward_workspace_bytes(B, P, F, ctypes.byref(n)),
                      "crdpn_pointnet_backward_workspace_bytes")
        ws_owner, wptr = _aligned(n.value, dev)
        grads = [torch.empty_like(t) for t in p]
        c1w, c1b, c2w, c2b, c3w, c3b, g1, b1, g2, b2, g3, b3 = p
        sync, total, px = ctx.sync
        bargs = (x.data_ptr(), B, P, F, c1w.data_ptr(), c2w.data_ptr(), c3w.data_ptr(),
                 g1.data_ptr(), b1.data_ptr(), g2.data_ptr(), b2.data_ptr(), g3.data_ptr(), b3.data_ptr(),
                 g.data_ptr(), cptr, cbytes, *[t.data_ptr() for t in grads], wptr, n.value)
        with _native.on_device(dev):
            if sync is None:
                rc = lib.crdpn_pointnet_backward(*bargs, _native.stream_ptr(dev))
                _native.check(rc, "crdpn_pointnet_backward")
            else:
                for ph in range(4):
                    rc = lib.crdpn_pointnet_backward_phased(*bargs, ph, ph + 1, total, _native.stream_ptr(dev))
                    _native.check(rc, "crdpn_pointnet_backward_phased")
                    if ph < 3:
                        _sum_over_ranks(lib, (B, P, F), 3 + ph, sync[0],
                                        {1: (ws_owner, wptr), 2: (grads[10], grads[10].data_ptr()), 3: (grads[11], grads[11].data_ptr())}, px)
        del owner, ws_owner
        return (None, None, *grads)
