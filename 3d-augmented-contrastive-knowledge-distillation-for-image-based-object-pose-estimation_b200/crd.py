"""CRD memory-bank NCE loss -- host-side mirror of the published CRD module surface, on B200 kernels.

The reference repository calls this path ``--crd`` (``trainingKD.py:127,282-283`` ->
``KD/common/base_class.py:303-436``) but ships no memory-bank code; the surface kept here
is the one ``BASELINE.json`` names and the published CRD algorithm defines:

    CRDLoss(opt).forward(f_s, f_t, idx, contrast_idx=None) -> loss
    Embed, Normalize, ContrastMemory, ContrastLoss, AliasMethod

with the same constructor arguments, parameter / buffer names (``embed_s.linear.weight``,
``contrast.memory_v1``, ``contrast.params`` ...) and semantics, so that it drops into the KD loop next to
``representation_loss`` (``KD/vision/vanilla/vanilla_kd.py:158-160``; call at ``base_class.py:387``).

All heavy work goes through the C ABI of ``libcrdpn_b200.so`` (``include/crdpn_b200.h``): one fused
gather-dot-exp-loss-backward pass, one deterministic reduction, one momentum-update launch.  There is no
CPU path: tensors must live on a CUDA device.
"""
from __future__ import annotations

import ctypes
import math
from types import SimpleNamespace

import torch
from torch import nn

from . import _native

EPS = 1e-7


_stream_ptr = _native.stream_ptr


def _require_cuda(t: torch.Tensor, name: str) -> None:
    if not t.is_cuda:
        raise RuntimeError(f"{name} must be a CUDA tensor: this package has no CPU fallback")


# ----------------------------------------------------------------------------------------------------------
class AliasMethod:
    """Alias-method sampler (published CRD ``AliasMethod``): ``draw(N)`` returns N int64 indices.

    Tables are built on the host by ``crdpn_alias_build`` (Vose stack pairing, fp32); draws run on the GPU
    with a counter-based Philox stream keyed by ``seed`` so they are reproducible and bit-exact against the
    oracle.  ``offset`` advances by the number of values drawn.
    """

    def __init__(self, probs: torch.Tensor, seed: int | None = None):
        probs = probs.detach().to("cpu", torch.float32).contiguous()  # normalised inside crdpn_alias_build
        n = probs.numel()
        self.prob = torch.zeros(n, dtype=torch.float32)
        self.alias = torch.zeros(n, dtype=torch.int64)
        _native.check(_native.lib().crdpn_alias_build(probs.data_ptr(), n, self.prob.data_ptr(),
                                                      self.alias.data_ptr()), "crdpn_alias_build")
        self.seed = int(torch.initial_seed() if seed is None else seed) & 0xFFFFFFFFFFFFFFFF
        self.offset = 0
        # uniform unigrams build prob == 1 everywhere: the draw then never reads the tables (same indices, no gather)
        self.uniform = bool((self.prob == 1.0).all())

    def table_ptrs(self):
        return (None, None) if self.uniform else (self.prob.data_ptr(), self.alias.data_ptr())

    def cuda(self, device=None):
        self.prob = self.prob.cuda(device)
        self.alias = self.alias.cuda(device)
        return self

    def to(self, device):
        self.prob = self.prob.to(device)
        self.alias = self.alias.to(device)
        return self

    def draw(self, N: int) -> torch.Tensor:
        _require_cuda(self.prob, "AliasMethod tables (call .cuda() first)")
        out = torch.empty(N, dtype=torch.int64, device=self.prob.device)
        with _native.on_device(self.prob.device):
            _native.check(_native.lib().crdpn_alias_draw(*self.table_ptrs(),
                                                         self.prob.numel(), N, self.seed, self.offset,
                                                         out.data_ptr(), _stream_ptr(self.prob.device)),
                          "crdpn_alias_draw")
        self.offset += N
        return out

    def draw_contrast(self, y: torch.Tensor, K1: int, row_base: int = 0) -> torch.Tensor:
        """[B, K1] contrast indices with column 0 = y (ContrastMemory.forward when idx is None); `row_base` is added to
        the drawn columns (a shard's sampler draws local rows, the contrast list holds global ones)."""
        _require_cuda(self.prob, "AliasMethod tables (call .cuda() first)")
        y = y.contiguous()
        B = y.numel()
        out = torch.empty(B, K1, dtype=torch.int64, device=self.prob.device)
        with _native.on_device(self.prob.device):
            _native.check(_native.lib().crdpn_alias_draw_contrast_local(*self.table_ptrs(), self.prob.numel(), int(row_base),
                                                                        y.data_ptr(), B, K1, self.seed, self.offset,
                                                                        out.data_ptr(), _stream_ptr(self.prob.device)),
                          "crdpn_alias_draw_contrast_local")
        self.offset += B * K1
        return out


# ----------------------------------------------------------------------------------------------------------
class Normalize(nn.Module):
    """x / ||x||_p along dim 1 (no epsilon), as in the published Embed tail."""

    def __init__(self, power: int = 2):
        super().__init__()
        self.power = power

    def forward(self, x):
        norm = x.pow(self.power).sum(1, keepdim=True).pow(1.0 / self.power)
        return x.div(norm)


class _EmbedFunction(torch.autograd.Function):
    """Linear + L2 normalise in two launches (crdpn_embed_forward); backward in two or three (crdpn_embed_backward)."""

    @staticmethod
    def forward(ctx, x, weight, bias):
        x, weight, bias = x.contiguous(), weight.contiguous(), bias.contiguous()
        B, dim_in = x.shape
        D = weight.shape[0]
        dev = x.device
        pre = torch.empty(B, D, dtype=torch.float32, device=dev)
        v = torch.empty_like(pre)
        inv = torch.empty(B, dtype=torch.float32, device=dev)
        with _native.on_device(dev):
            rc = _native.lib().crdpn_embed_forward(x.data_ptr(), weight.data_ptr(), bias.data_ptr(), B, dim_in, D,
                                                   pre.data_ptr(), v.data_ptr(), inv.data_ptr(), _stream_ptr(dev))
        _native.check(rc, "crdpn_embed_forward")
        ctx.save_for_backward(x, weight, v, inv)
        ctx.need_dx = ctx.needs_input_grad[0]
        return v

    @staticmethod
    def backward(ctx, grad_v):
        x, weight, v, inv = ctx.saved_tensors
        B, dim_in = x.shape
        D = weight.shape[0]
        dev = x.device
        g = grad_v.contiguous()
        dW = torch.empty_like(weight)
        db = torch.empty(D, dtype=torch.float32, device=dev)
        dx = torch.empty_like(x) if ctx.need_dx else None
        d_pre = torch.empty(B, D, dtype=torch.float32, device=dev)
        with _native.on_device(dev):
            rc = _native.lib().crdpn_embed_backward(x.data_ptr(), weight.data_ptr(), v.data_ptr(), inv.data_ptr(),
                                                    g.data_ptr(), None, B, dim_in, D, dW.data_ptr(), db.data_ptr(),
                                                    dx.data_ptr() if dx is not None else None, d_pre.data_ptr(),
                                                    _stream_ptr(dev))
        _native.check(rc, "crdpn_embed_backward")
        return dx, dW, db


class _CRDLossFunction(torch.autograd.Function):
    """The whole CRD step as one autograd node and TWO foreign calls: ``crdpn_crd_loss_forward`` (embed heads ->
    negative draw -> fused score/loss/backward -> reduction + momentum update, 5 launches) and
    ``crdpn_crd_loss_backward`` (both embed-head backwards, 2 launches).  Everything the step produces lives in one
    float32 arena ``[result(16) | pre_s | pre_t | v1 | v2 | grad_v1 | grad_v2 | inv1 | inv2]``; the upstream gradient
    of the loss is a device scalar folded into the embed backward, so no elementwise torch kernels run at all."""

    @staticmethod
    def forward(ctx, f_s, f_t, Ws, bs, Wt, bt, y, contrast_idx, crit):
        mem = crit.contrast
        dev = f_s.device
        xs, xt = f_s.detach(), f_t.detach()
        if xs.dim() != 2 or not xs.is_contiguous():
            xs = xs.reshape(xs.shape[0], -1).contiguous()
        if xt.dim() != 2 or not xt.is_contiguous():
            xt = xt.reshape(xt.shape[0], -1).contiguous()
        Wsc, bsc, Wtc, btc = (t.detach() if t.is_contiguous() else t.detach().contiguous() for t in (Ws, bs, Wt, bt))
        B, D = xs.shape[0], Wsc.shape[0]
        hp = mem._host_params()
        K1 = hp.K + 1
        if xt.shape[0] != B or y.numel() != B:
            raise RuntimeError("f_s, f_t and idx must share the batch dimension")
        mem._check_device(y, "idx")
        yc = y if (y.dtype == torch.int64 and y.is_contiguous()) else y.contiguous().to(torch.int64)
        if contrast_idx is not None:
            mem._check_device(contrast_idx, "contrast_idx")
            if contrast_idx.dtype not in (torch.int64, torch.int32):
                contrast_idx = contrast_idx.to(torch.int64)
            contrast_idx = contrast_idx.contiguous()   # int32 lists are consumed as they are (half the bytes to copy and scan)
            if contrast_idx.shape != (B, K1):
                raise RuntimeError(f"contrast_idx must have shape [B, K+1] = {(B, K1)}, got {tuple(contrast_idx.shape)}")
        BD = B * D
        Bp = (B + 3) & ~3
        arena = torch.empty(16 + 6 * BD + 2 * Bp, dtype=torch.float32, device=dev)
        # the 8 result doubles live in their own tensor: the returned loss is a view of it, and it is NOT saved for
        # backward, so an in-place op on the loss (loss /= accum) cannot invalidate the saved arena
        result = torch.empty(16, dtype=torch.float32, device=dev)
        base = arena.data_ptr()
        o_pre_s, o_pre_t, o_v1, o_v2, o_g1, o_g2 = (base + 4 * (16 + i * BD) for i in range(6))
        o_inv1 = base + 4 * (16 + 6 * BD)
        o_inv2 = o_inv1 + 4 * Bp
        if hp.Z1 <= 0 or hp.Z2 <= 0:
            # first call: the published algorithm freezes Z from this batch (one device->host read); run the embed
            # heads on their own, freeze, then take the fused path below with the SAME drawn negatives
            lib = _native.lib()
            with _native.on_device(dev):
                st = _stream_ptr(dev)
                _native.check(lib.crdpn_embed_forward(xs.data_ptr(), Wsc.data_ptr(), bsc.data_ptr(), B, xs.shape[1], D,
                                                      o_pre_s, o_v1, o_inv1, st), "crdpn_embed_forward")
                _native.check(lib.crdpn_embed_forward(xt.data_ptr(), Wtc.data_ptr(), btc.data_ptr(), B, xt.shape[1], D,
                                                      o_pre_t, o_v2, o_inv2, st), "crdpn_embed_forward")
            v1 = arena[16 + 2 * BD:16 + 3 * BD].view(B, D)
            v2 = arena[16 + 3 * BD:16 + 4 * BD].view(B, D)
            if contrast_idx is None:
                contrast_idx = mem.multinomial.draw_contrast(yc, K1)
            mem._freeze_z(v1, v2, contrast_idx.to(torch.int64))
            hp = mem._host_params()
        m1, m2, stride, dt = mem._banks()
        variant = mem._step_variant(B, K1, D)
        ws = mem._workspace(B, K1, D, dev, variant)
        if contrast_idx is not None and contrast_idx.dtype == torch.int32:
            if variant & mem.STREAM:
                contrast_idx = contrast_idx.to(torch.int64)   # the bucketing passes of the streaming kernels read int64
            else:
                variant |= mem.IDX32
        smp = mem.multinomial
        dev_offset = None
        if contrast_idx is None and smp.uniform and not (variant & mem.STREAM):
            cidx_ptr, scratch_ptr = None, None     # the scoring pass draws the negatives itself: no list in memory
            dev_offset = mem._device_offset(dev)
            if dev_offset is not None:             # CUDA-graph replays: the advancing part of the offset lives on the device
                variant |= mem.DEVICE_OFFSET
                scratch_ptr = dev_offset.data_ptr()
        elif contrast_idx is None:
            scratch = mem._idx_scratch
            if scratch is None or scratch.numel() != B * K1 or scratch.device != dev:
                scratch = mem._idx_scratch = torch.empty(B * K1, dtype=torch.int64, device=dev)
            cidx_ptr, scratch_ptr = None, scratch.data_ptr()
        else:
            cidx_ptr, scratch_ptr = contrast_idx.data_ptr(), None
        with _native.on_device(dev):
            rc = _native.lib().crdpn_crd_loss_forward(
                xs.data_ptr(), xs.shape[1], Wsc.data_ptr(), bsc.data_ptr(),
                xt.data_ptr(), xt.shape[1], Wtc.data_ptr(), btc.data_ptr(),
                yc.data_ptr(), cidx_ptr, *smp.table_ptrs(), smp.seed, smp.offset, scratch_ptr,
                m1.data_ptr(), m2.data_ptr(), stride, dt,
                B, K1, D, mem.nLem, mem.k_total, mem.row_begin, mem.row_end,
                hp.T, hp.Z1, hp.Z2, EPS, hp.m, 1.0 - hp.m,
                o_pre_s, o_pre_t, o_v1, o_v2, o_inv1, o_inv2,
                result.data_ptr(), o_g1, o_g2, ws.data_ptr(), ws.numel(), variant, _stream_ptr(dev))
        _native.check(rc, "crdpn_crd_loss_forward")
        if contrast_idx is None and dev_offset is None:
            smp.offset += B * K1
        ctx.save_for_backward(arena, xs, xt, Wsc, Wtc)
        ctx.need_dx = (ctx.needs_input_grad[0], ctx.needs_input_grad[1])
        ctx.in_shapes = (f_s.shape, f_t.shape)
        return result[12]  # float32(loss_s + loss_t), written by the reduction kernel into result slot 6

    @staticmethod
    def backward(ctx, grad_out):
        arena, xs, xt, Ws, Wt = ctx.saved_tensors
        dev = xs.device
        B, D = xs.shape[0], Ws.shape[0]
        BD = B * D
        Bp = (B + 3) & ~3
        base = arena.data_ptr()
        o_v1, o_v2, o_g1, o_g2 = (base + 4 * (16 + i * BD) for i in range(2, 6))
        o_inv1 = base + 4 * (16 + 6 * BD)
        o_inv2 = o_inv1 + 4 * Bp
        scale = grad_out.detach().to(torch.float32).contiguous()
        dWs, dWt = torch.empty_like(Ws), torch.empty_like(Wt)
        dbs = torch.empty(2 * D, dtype=torch.float32, device=dev)
        dxs = torch.empty_like(xs) if ctx.need_dx[0] else None
        dxt = torch.empty_like(xt) if ctx.need_dx[1] else None
        d_pre = torch.empty(2 * BD, dtype=torch.float32, device=dev)
        with _native.on_device(dev):
            rc = _native.lib().crdpn_crd_loss_backward(
                xs.data_ptr(), xs.shape[1], Ws.data_ptr(), o_v1, o_inv1, o_g1,
                xt.data_ptr(), xt.shape[1], Wt.data_ptr(), o_v2, o_inv2, o_g2,
                scale.data_ptr(), B, D,
                dWs.data_ptr(), dbs.data_ptr(), dxs.data_ptr() if dxs is not None else None,
                dWt.data_ptr(), dbs.data_ptr() + 4 * D, dxt.data_ptr() if dxt is not None else None,
                d_pre.data_ptr(), _stream_ptr(dev))
        _native.check(rc, "crdpn_crd_loss_backward")
        if dxs is not None:
            dxs = dxs.view(ctx.in_shapes[0])
        if dxt is not None:
            dxt = dxt.view(ctx.in_shapes[1])
        return dxs, dxt, dWs, dbs[:D], dWt, dbs[D:], None, None, None


class Embed(nn.Module):
    """flatten -> Linear(dim_in, dim_out) -> L2 normalise.

    CUDA float32 inputs run the fused kernels; the module keeps the published sub-module names (``linear``,
    ``l2norm``) so state_dicts and optimiser parameter groups are unchanged.  There is no eager / CPU path: anything
    but a CUDA float32 input raises (the gloo host-logic tests substitute their own subclass)."""

    def __init__(self, dim_in: int = 1024, dim_out: int = 128):
        super().__init__()
        self.linear = nn.Linear(dim_in, dim_out)
        self.l2norm = Normalize(2)

    def forward(self, x):
        x = x.view(x.shape[0], -1)
        _require_cuda(x, "Embed input")
        if x.dtype != torch.float32 or self.linear.weight.dtype != torch.float32:
            raise RuntimeError("Embed runs in float32 only (input and Linear parameters)")
        return _EmbedFunction.apply(x, self.linear.weight, self.linear.bias)


class ContrastLoss(nn.Module):
    """NCE criterion on normalised scores x[B, K+1, 1] (column 0 positive); unfused reference form.

    CRDLoss does not call this (the fused kernel computes the same sum); it is kept for API parity and
    for users who call ContrastMemory directly."""

    def __init__(self, n_data: int):
        super().__init__()
        self.n_data = n_data

    def forward(self, x):
        bsz, m = x.shape[0], x.size(1) - 1
        Pn = 1.0 / float(self.n_data)
        P_pos = x.select(1, 0)
        log_D1 = torch.div(P_pos, P_pos.add(m * Pn + EPS)).log_()
        P_neg = x.narrow(1, 1, m)
        log_D0 = torch.div(P_neg.clone().fill_(m * Pn), P_neg.add(m * Pn + EPS)).log_()
        return -(log_D1.sum(0) + log_D0.view(-1, 1).sum(0)) / bsz


# ----------------------------------------------------------------------------------------------------------
class _FusedCRDFunction(torch.autograd.Function):
    """loss = NCE(out_v1) + NCE(out_v2); gradients w.r.t. v1, v2 come out of the same kernel pass."""

    @staticmethod
    def forward(ctx, v1, v2, y, contrast_idx, mem):
        loss, g1, g2 = mem._score_and_update(v1.detach(), v2.detach(), y, contrast_idx)
        ctx.save_for_backward(g1, g2)
        return loss

    @staticmethod
    def backward(ctx, grad_out):
        g1, g2 = ctx.saved_tensors
        return grad_out * g1, grad_out * g2, None, None, None


class _ContrastOutFunction(torch.autograd.Function):
    """The unfused published surface, differentiable: (v1, v2) -> (out_v1, out_v2) [B, K+1, 1] with the momentum update
    as a side effect; backward = crdpn_crd_out_backward (gradients through the PRE-update rows, as the published code's
    detached copy gives)."""

    @staticmethod
    def forward(ctx, v1, v2, y, idx, mem):
        v1c, v2c = v1.detach(), v2.detach()
        mem._freeze_z(v1c, v2c, idx)
        hp = mem._host_params()
        m1, m2, _, _ = mem._banks()
        rows = mem.row_end - mem.row_begin
        yl = (y - mem.row_begin).clamp_(0, max(rows - 1, 0))          # rows this shard does not own are never read back
        old1, old2 = m1.index_select(0, yl).float().contiguous(), m2.index_select(0, yl).float().contiguous()
        _, _, _, o1, o2 = mem._score(v1c, v2c, idx, hp.Z1, hp.Z2, want_out=True)
        mem._update(v1c, v2c, y)
        ctx.save_for_backward(old1, old2, y, idx, o1, o2)
        ctx.mem = mem
        return o1.unsqueeze(-1), o2.unsqueeze(-1)

    @staticmethod
    def backward(ctx, go1, go2):
        old1, old2, y, idx, o1, o2 = ctx.saved_tensors
        mem = ctx.mem
        B, K1 = idx.shape
        D = old1.shape[1]
        dev = old1.device
        m1, m2, stride, dt = mem._banks()
        go1 = (torch.zeros_like(o1) if go1 is None else go1.reshape(B, K1).to(torch.float32)).contiguous()
        go2 = (torch.zeros_like(o2) if go2 is None else go2.reshape(B, K1).to(torch.float32)).contiguous()
        n = ctypes.c_size_t(0)
        lib = _native.lib()
        _native.check(lib.crdpn_crd_out_backward_workspace_bytes(B, K1, D, ctypes.byref(n)), "crdpn_crd_out_backward_workspace_bytes")
        ws = torch.empty(n.value, dtype=torch.uint8, device=dev)
        g1 = torch.empty(B, D, dtype=torch.float32, device=dev)
        g2 = torch.empty(B, D, dtype=torch.float32, device=dev)
        with _native.on_device(dev):
            rc = lib.crdpn_crd_out_backward(m1.data_ptr(), m2.data_ptr(), stride, dt, old1.data_ptr(), old2.data_ptr(),
                                            y.data_ptr(), idx.data_ptr(), go1.data_ptr(), go2.data_ptr(), o1.data_ptr(),
                                            o2.data_ptr(), B, K1, D, mem.row_begin, mem.row_end, mem._host_params().T,
                                            g1.data_ptr(), g2.data_ptr(), ws.data_ptr(), ws.numel(), _stream_ptr(dev))
        _native.check(rc, "crdpn_crd_out_backward")
        return g1, g2, None, None, None


class ContrastMemory(nn.Module):
    """Two momentum memory banks with NCE scoring (published CRD ``ContrastMemory``).

    Buffers (state_dict-compatible): ``params`` = [K, T, Z_v1, Z_v2, momentum], ``memory_v1``,
    ``memory_v2`` of shape [n_rows, inputSize].  HBM layout: both banks live interleaved in one
    [n_rows, 2, inputSize] allocation (``memory_v1``/``memory_v2`` are strided views of it) so one sampled
    index touches one contiguous span; the kernels accept any common row pitch, so separately allocated
    banks keep working.

    Sharding: with ``row_begin/row_end`` set, this rank holds rows [row_begin,row_end) of a global
    ``outputSize``-row bank and scores / updates only those (see sharded.py for the collectives).
    """

    def __init__(self, inputSize: int, outputSize: int, K: int, T: float = 0.07, momentum: float = 0.5,
                 row_begin: int = 0, row_end: int | None = None, bank_dtype: torch.dtype = torch.float32,
                 interleave: bool = True, seed: int | None = None):
        super().__init__()
        self.nLem = outputSize
        self.row_begin = int(row_begin)
        self.row_end = int(outputSize if row_end is None else row_end)
        self.interleave = interleave
        self.unigrams = torch.ones(self.nLem)
        self.multinomial = AliasMethod(self.unigrams, seed=seed)
        self.K = K
        self.k_total = 0   # negatives per anchor over all shards (0: the K+1 columns of contrast_idx are all of them)
        self.variant = 0
        self.streaming = None    # None: automatic (bf16 banks only), True / False: forced (see _step_variant)
        self.sweep = None        # None: automatic, True / False: forced (band-sorted lists, see _step_variant)
        self.register_buffer("params", torch.tensor([K, T, -1, -1, momentum], dtype=torch.float32))
        stdv = 1.0 / math.sqrt(inputSize / 3)
        rows = self.row_end - self.row_begin
        m1 = torch.rand(rows, inputSize).mul_(2 * stdv).add_(-stdv).to(bank_dtype)
        m2 = torch.rand(rows, inputSize).mul_(2 * stdv).add_(-stdv).to(bank_dtype)
        self.register_buffer("memory_v1", m1)
        self.register_buffer("memory_v2", m2)
        self._relayout()
        self._ws = None
        self._ws_cache = {}
        self._res = None
        self._idx_scratch = None
        self._dev_offset = None   # see device_sampler_offset
        # host mirror of params (avoids a device->host read per step once Z is frozen)
        self._host = None

    # -- sampler offset on the device (CUDA-graph replays) -----------------------------------------------------
    def device_sampler_offset(self, enabled: bool = True):
        """Keep the ADVANCING part of the negative sampler's Philox offset in device memory: the scoring pass adds the
        counter to the (then fixed) host offset and a one-thread kernel advances it by B * (K+1) after every step.  A step
        captured in a CUDA graph (``GraphedStep``) then draws fresh negatives on every replay -- the same negatives the
        eager loop would have drawn -- instead of replaying the offset that was baked in at capture.  Applies to the
        in-kernel uniform draw (``CRDLoss(f_s, f_t, idx)`` with uniform unigrams).  Disabling folds the counter back into
        the host offset (one device read)."""
        if enabled and self._dev_offset is None:
            self._dev_offset = "pending"      # allocated on the step's device at first use
        elif not enabled and self._dev_offset is not None:
            if isinstance(self._dev_offset, torch.Tensor):
                smp = self._sampler_for_offset()
                smp.offset += int(self._dev_offset.item())
            self._dev_offset = None
        return self

    def _sampler_for_offset(self):
        return self.multinomial

    def sampler_state(self) -> dict:
        """(seed, offset) of the negative sampler's Philox stream.  Not part of ``state_dict`` (its keys stay those of the
        published module); save it next to a checkpoint and hand it to ``load_sampler_state`` to resume the SAME stream of
        negatives instead of replaying it from offset 0."""
        smp = self._sampler_for_offset()
        extra = int(self._dev_offset.item()) if isinstance(self._dev_offset, torch.Tensor) else 0
        return {"seed": int(smp.seed), "offset": int(smp.offset) + extra}

    def load_sampler_state(self, state: dict):
        smp = self._sampler_for_offset()
        smp.seed, smp.offset = int(state["seed"]), int(state["offset"])
        if isinstance(self._dev_offset, torch.Tensor):
            self._dev_offset.zero_()
        return self

    def _device_offset(self, device):
        if self._dev_offset is None:
            return None
        if not isinstance(self._dev_offset, torch.Tensor) or self._dev_offset.device != device:
            self._dev_offset = torch.zeros(1, dtype=torch.int64, device=device)
        return self._dev_offset

    # -- layout ---------------------------------------------------------------------------------------
    def _relayout(self):
        m1, m2 = self._buffers["memory_v1"], self._buffers["memory_v2"]
        if self.interleave:
            bank = torch.stack([m1, m2], dim=1).contiguous()  # [rows, 2, D]
            self._buffers["memory_v1"] = bank[:, 0, :]
            self._buffers["memory_v2"] = bank[:, 1, :]
        else:
            self._buffers["memory_v1"] = m1.contiguous()
            self._buffers["memory_v2"] = m2.contiguous()

    def _apply(self, fn, *args, **kwargs):
        super()._apply(fn, *args, **kwargs)
        self._relayout()  # .cuda()/.to() de-interleave the two views; put them back in one allocation
        self.multinomial.to(self._buffers["memory_v1"].device)
        self._ws = None
        self._ws_cache = {}
        self._idx_scratch = None
        self._host = None
        return self

    def _load_from_state_dict(self, *args, **kwargs):
        super()._load_from_state_dict(*args, **kwargs)
        self._host = None

    def _banks(self):
        m1, m2 = self._buffers["memory_v1"], self._buffers["memory_v2"]
        if m1.stride(1) != 1 or m2.stride(1) != 1 or m1.stride(0) != m2.stride(0) or m1.dtype != m2.dtype:
            self._relayout()
            m1, m2 = self.memory_v1, self.memory_v2
        dt = _native.F32 if m1.dtype == torch.float32 else _native.BF16
        if m1.dtype not in (torch.float32, torch.bfloat16):
            raise RuntimeError("memory banks must be float32 or bfloat16")
        return m1, m2, m1.stride(0), dt

    def _host_params(self):
        if self._host is None:
            p = self.params.detach().cpu().tolist()
            self._host = SimpleNamespace(K=int(p[0]), T=float(p[1]), Z1=float(p[2]), Z2=float(p[3]), m=float(p[4]))
        return self._host

    STREAM = 0x200   # variant bit: bank-streaming formulation of the step (csrc/crd_stream.cuh)
    IDX32 = 0x1000   # variant bit: contrast_idx is an int32 list
    SWEEP = 0x400    # variant bit: band-sorted contrast lists (csrc/crd_kernels.cu crd_band_sort_kernel): repeats of a row are L2 hits
    DEVICE_OFFSET = 0x4000   # variant bit: the sampler offset's advancing part is a device counter (CUDA-graph replays)

    def _step_variant(self, B, K1, D):
        """Variant passed to crdpn_crd_step.  ``self.streaming`` selects the bank-streaming formulation, which reads every
        resident row once instead of gathering B*(K+1) rows:

        * ``None`` (default) -- automatic: used for bf16 banks (the tcgen05 tensor-core kernel of csrc/crd_tc_stream.cuh,
          1.8x faster than the bf16 gather kernel at the headline shape) when feat_dim is 128, the batch is <= 48 and the
          step draws at least two samples per resident row; never for fp32 banks.
        * ``True`` -- forced (fp32 banks: the EXPERIMENTAL register kernel of csrc/crd_stream.cuh, instruction-bound and
          2x slower than the gather kernel); raises if the shape is not supported.
        * ``False`` -- the gather kernel."""
        if self.variant & self.STREAM:
            return self.variant
        if self.streaming is False:
            return self._with_sweep(self.variant, B, K1, D)
        rows = self.row_end - self.row_begin
        ok = D == 128 and 1 <= B <= 48 and rows >= 1
        if self.streaming is None:
            # samples that land on this shard: all K+1 columns when every rank draws its own rows (k_total > 0), else its share
            hits = B * K1 if self.k_total > 0 else B * K1 * rows // max(self.nLem, 1)
            auto = ok and self._buffers["memory_v1"].dtype == torch.bfloat16 and hits >= 2 * rows
            return self.variant | self.STREAM if auto else self._with_sweep(self.variant, B, K1, D)
        if not ok:
            raise RuntimeError("streaming CRD step needs feat_dim 128 and batch <= 48")
        return self.variant | self.STREAM

    def _with_sweep(self, variant, B, K1, D):
        """Adds the SWEEP bit to a gather-kernel variant when the band-sorted formulation pays: the step draws each resident
        row more than ~1.5 times and the shard is about as large as the 126 MB L2 or larger (a bank well inside the L2 -- config 0,
        92 MB -- has its repeats as L2 hits already and the pre-pass only costs).  Measured at the headline shape:
        DRAM traffic 2.91 -> 1.02 GB, scoring kernel 0.433 -> 0.296 ms, ~12 us of pre-pass; 2-way / 4-way / 8-way shards 0.227 -> 0.179 /
        0.122 -> 0.110 / 0.079 -> 0.076 ms per step (profiles/r2_sweep_ab.py).  ``self.sweep`` forces it on / off."""
        if variant & (self.SWEEP | 0x40 | 0x100):
            return variant
        if self.sweep is False:
            return variant
        if self.sweep is None:
            rows = self.row_end - self.row_begin
            esz = 2 if self._buffers["memory_v1"].dtype == torch.bfloat16 else 4
            hits = B * K1 if self.k_total > 0 else B * K1 * rows // max(self.nLem, 1)
            # (bf16 banks: measured no gain, 0.287 -> 0.300 ms at the headline shape -- their gather is transaction-bound)
            if not (esz == 4 and rows * 2 * D * esz >= (120 << 20) and 2 * hits >= 3 * rows and B * K1 >= (1 << 20)):
                return variant
        return variant | self.SWEEP

    def _workspace(self, B, K1, D, device, variant=0):
        stream = bool(variant & self.STREAM)
        key = (B, K1, D, device, stream)
        cache = self._ws_cache
        if cache.get(stream, (None, None))[0] != key:
            n = ctypes.c_size_t(0)
            if stream:
                _native.check(_native.lib().crdpn_crd_stream_workspace_bytes(B, K1, D, self.row_end - self.row_begin,
                                                                            device.index or 0, ctypes.byref(n)),
                              "crdpn_crd_stream_workspace_bytes")
            else:
                _native.check(_native.lib().crdpn_crd_workspace_bytes(B, K1, D, device.index or 0, ctypes.byref(n)),
                              "crdpn_crd_workspace_bytes")
            cache[stream] = (key, torch.empty(n.value, dtype=torch.uint8, device=device))
        self._ws = cache[stream][1]
        if self._res is None or self._res.device != device:
            self._res = torch.zeros(8, dtype=torch.float64, device=device)
        return self._ws

    # -- kernel calls ---------------------------------------------------------------------------------
    def _score(self, v1, v2, idx, Z1, Z2, want_out=False, result=None):
        """One crdpn_crd_score call. Returns (result[8] f64 device tensor, grad_v1, grad_v2, out_v1, out_v2)."""
        m1, m2, stride, dt = self._banks()
        B, K1 = idx.shape
        D = v1.shape[1]
        dev = v1.device
        ws = self._workspace(B, K1, D, dev)
        res = self._res if result is None else result
        full = Z1 > 0 and Z2 > 0
        g1 = torch.empty_like(v1) if full else None
        g2 = torch.empty_like(v2) if full else None
        o1 = torch.empty(B, K1, dtype=torch.float32, device=dev) if want_out else None
        o2 = torch.empty(B, K1, dtype=torch.float32, device=dev) if want_out else None
        hp = self._host_params()
        with _native.on_device(dev):
            rc = _native.lib().crdpn_crd_score(
                m1.data_ptr(), m2.data_ptr(), stride, dt, v1.data_ptr(), v2.data_ptr(), idx.data_ptr(),
                B, K1, D, self.nLem, self.k_total, self.row_begin, self.row_end,
                hp.T, Z1, Z2, EPS,
                o1.data_ptr() if want_out else None, o2.data_ptr() if want_out else None,
                res.data_ptr(), g1.data_ptr() if full else None, g2.data_ptr() if full else None,
                ws.data_ptr(), ws.numel(), self.variant, _stream_ptr(dev))
        _native.check(rc, "crdpn_crd_score")
        return res, g1, g2, o1, o2

    def _update(self, v1, v2, y):
        m1, m2, stride, dt = self._banks()
        hp = self._host_params()
        m32 = hp.m  # read back from the fp32 `params` buffer, so already an exact fp32 value
        with _native.on_device(v1.device):
            rc = _native.lib().crdpn_crd_momentum_update(
                m1.data_ptr(), m2.data_ptr(), stride, dt, v1.data_ptr(), v2.data_ptr(), y.data_ptr(),
                v1.shape[0], v1.shape[1], self.row_begin, self.row_end, m32, 1.0 - m32, _stream_ptr(v1.device))
        _native.check(rc, "crdpn_crd_momentum_update")

    def _step(self, v1, v2, y, idx, Z1, Z2):
        """crdpn_crd_step: score + loss + backward, then reduction and momentum update in one launch."""
        m1, m2, stride, dt = self._banks()
        B, K1 = idx.shape
        D = v1.shape[1]
        dev = v1.device
        variant = self._step_variant(B, K1, D)
        ws = self._workspace(B, K1, D, dev, variant)
        res = torch.empty(8, dtype=torch.float64, device=dev)  # fresh: the returned loss is a view into it
        g1, g2 = self._grad_buffers(v1, v2)
        hp = self._host_params()
        with _native.on_device(dev):
            rc = _native.lib().crdpn_crd_step(
                m1.data_ptr(), m2.data_ptr(), stride, dt, v1.data_ptr(), v2.data_ptr(), idx.data_ptr(), y.data_ptr(),
                B, K1, D, self.nLem, self.k_total, self.row_begin, self.row_end, hp.T, Z1, Z2, EPS, hp.m, 1.0 - hp.m,
                res.data_ptr(), g1.data_ptr(), g2.data_ptr(), ws.data_ptr(), ws.numel(), variant, _stream_ptr(dev))
        _native.check(rc, "crdpn_crd_step")
        return res, g1, g2

    def _grad_buffers(self, v1, v2):
        """Where the kernel writes grad_v1 / grad_v2 (the sharded subclass hands out slices of its packed
        all-reduce buffer so nothing has to be copied before the exchange)."""
        return torch.empty_like(v1), torch.empty_like(v2)

    def _reduce_sums(self, res):
        """Hook for the sharded subclass: sum the first-call (sum_e1, sum_e2, count) over ranks."""
        return res

    def _reduce_partials(self, res, g1, g2):
        """Hook for the sharded subclass: sum loss partials and grad_v over ranks."""
        return res, g1, g2

    def _check_device(self, t, name):
        _require_cuda(t, name)

    def _prepare(self, v1, v2, y, idx):
        for t, name in ((v1, "v1"), (v2, "v2"), (y, "y")):
            self._check_device(t, name)
        if v1.dtype != torch.float32 or v2.dtype != torch.float32:
            raise RuntimeError("embeddings must be float32")
        v1, v2, y = v1.contiguous(), v2.contiguous(), y.contiguous().to(torch.int64)
        K1 = self._host_params().K + 1
        if idx is None:
            idx = self.multinomial.draw_contrast(y, K1)
        else:
            self._check_device(idx, "contrast_idx")
            idx = idx.contiguous().to(torch.int64)
            if idx.shape != (v1.shape[0], K1):
                raise RuntimeError(f"contrast_idx must have shape [B, K+1] = {(v1.shape[0], K1)}, got {tuple(idx.shape)}")
        return v1, v2, y, idx

    def _freeze_z(self, v1, v2, idx):
        """First call: Z = mean(exp(s/T)) * n_data, stored as constants (one device->host read, as in the
        published algorithm)."""
        hp = self._host_params()
        if hp.Z1 > 0 and hp.Z2 > 0:
            return
        res, *_ = self._score(v1, v2, idx, -1.0, -1.0)
        res = self._reduce_sums(res.clone())
        s1, s2, cnt = res[2].item(), res[3].item(), res[4].item()
        if hp.Z1 <= 0:
            hp.Z1 = float(torch.tensor(s1 / cnt * self.nLem, dtype=torch.float32).item())
            self.params[2] = hp.Z1
            print("normalization constant Z_v1 is set to {:.1f}".format(hp.Z1))
        if hp.Z2 <= 0:
            hp.Z2 = float(torch.tensor(s2 / cnt * self.nLem, dtype=torch.float32).item())
            self.params[3] = hp.Z2
            print("normalization constant Z_v2 is set to {:.1f}".format(hp.Z2))

    def _score_and_update(self, v1, v2, y, idx):
        """Fused path used by CRDLoss: returns (loss 0-dim f32, grad_v1, grad_v2) and updates the banks."""
        self._freeze_z(v1, v2, idx)
        hp = self._host_params()
        res, g1, g2 = self._step(v1, v2, y, idx, hp.Z1, hp.Z2)
        red, g1, g2 = self._reduce_partials(res, g1, g2)
        if red is res:   # single shard: the kernel already wrote float32(loss_s + loss_t) into slot 6 -> no cast kernel
            loss = res.view(torch.float32)[12]
        else:
            loss = (red[0] + red[1]).to(torch.float32)
        return loss, g1, g2

    def fused_loss(self, v1, v2, y, idx=None):
        v1c, v2c, y, idx = self._prepare(v1, v2, y, idx)
        return _FusedCRDFunction.apply(v1c, v2c, y, idx, self)

    def forward(self, v1, v2, y, idx=None):
        """Published surface: returns (out_v1, out_v2), each [B, K+1, 1], and updates the banks.

        The outputs come out of the same fused scoring pass and are differentiable w.r.t. v1 / v2
        (``_ContrastOutFunction``), so ``ContrastLoss(out_v1) + ContrastLoss(out_v2)`` -- the published CRDLoss body --
        trains as written; ``CRDLoss`` / ``fused_loss`` are the fast path (one pass instead of two)."""
        v1c, v2c, y, idx = self._prepare(v1, v2, y, idx)
        return _ContrastOutFunction.apply(v1c, v2c, y, idx, self)


class CRDLoss(nn.Module):
    """CRD loss with two symmetric parts (published ``CRDLoss``).

    Args (``opt`` attributes): s_dim, t_dim, feat_dim, n_data, nce_k, nce_t, nce_m.
    forward(f_s [B,s_dim], f_t [B,t_dim], idx [B] int64, contrast_idx [B, nce_k+1] int64 or None) -> 0-dim loss.
    """

    def __init__(self, opt, **memory_kwargs):
        super().__init__()
        self.embed_s = Embed(opt.s_dim, opt.feat_dim)
        self.embed_t = Embed(opt.t_dim, opt.feat_dim)
        self.contrast = ContrastMemory(opt.feat_dim, opt.n_data, opt.nce_k, opt.nce_t, opt.nce_m, **memory_kwargs)
        self.criterion_t = ContrastLoss(opt.n_data)
        self.criterion_s = ContrastLoss(opt.n_data)

    def forward(self, f_s, f_t, idx, contrast_idx=None):
        _require_cuda(f_s, "f_s")
        _require_cuda(f_t, "f_t")
        if f_s.dtype != torch.float32 or f_t.dtype != torch.float32:
            raise RuntimeError("CRDLoss expects float32 features (cast f_s / f_t with .float() under autocast)")
        if type(self.contrast) is ContrastMemory:
            return _CRDLossFunction.apply(f_s, f_t, self.embed_s.linear.weight, self.embed_s.linear.bias,
                                          self.embed_t.linear.weight, self.embed_t.linear.bias, idx, contrast_idx, self)
        f_s = self.embed_s(f_s)   # a ContrastMemory subclass brings its own step: the per-op path
        f_t = self.embed_t(f_t)
        return self.contrast.fused_loss(f_s, f_t, idx, contrast_idx)
