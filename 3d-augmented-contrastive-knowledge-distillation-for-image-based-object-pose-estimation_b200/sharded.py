"""Row-sharded CRD memory banks over 1/2/4/8 GPUs of one box (one process per GPU, torch.distributed / NCCL).

The reference is single-process, single-GPU; this is the build's data-parallel
extension of the CRD step:

* both banks are sharded by sample index: rank r owns the contiguous rows ``shard_bounds(n_data, R, r)``;
* the batch is data parallel: every rank embeds its own ``B_loc`` anchors;
* exchange 1 (before the kernel): ONE all-gather of the packed ``(v1, v2, idx)`` rows -> all ``B`` anchors;
* every rank scores all ``B`` anchors against the negatives that live in ITS shard (``crdpn_crd_step`` with
  ``row_begin/row_end``; entries of other shards are dropped inside the kernel before any row is loaded) and
  momentum-updates the positive rows it owns;
* exchange 2 (after the kernel): ONE all-reduce of a packed fp32 buffer ``[grad_v1[B,D], grad_v2[B,D], loss_s,
  loss_t, ...]`` (about 47 KB at B=46, D=128) that the kernel's reduction writes in place.

Negatives: either a replicated ``contrast_idx[B, K+1]`` (parity mode: every rank scans the whole list and keeps
what it owns -- results equal the unsharded module up to fp32 summation order), or ``local_negatives=True``:
each rank draws / receives ``K_loc`` negatives inside its own shard (``contrast_idx[B, K_loc+1]`` per rank, column
0 still the global positive index, which non-owners skip), so no rank ever touches another rank's index list and
the per-rank bytes are exactly ``1/R`` of the total.  The NCE constant uses ``K = sum_r K_loc``.
"""
from __future__ import annotations

import torch
import torch.distributed as dist
from torch import nn

import ctypes

from . import _native
from .crd import ContrastMemory, Embed, _FusedCRDFunction, _stream_ptr


def shard_bounds(n: int, world: int, rank: int) -> tuple[int, int]:
    """Contiguous near-equal split of n rows: rank r owns [n*r//world, n*(r+1)//world)."""
    return n * rank // world, n * (rank + 1) // world


def pack_anchor_rows(v1: torch.Tensor, v2: torch.Tensor, y: torch.Tensor, rows: int) -> torch.Tensor:
    """[rows, 2D+2] fp32: v1 | v2 | y as two bit-cast fp32 words.  Rows beyond len(y) are padding (y = -1)."""
    b, d = v1.shape
    buf = torch.zeros(rows, 2 * d + 2, dtype=torch.float32, device=v1.device)
    buf[:b, :d] = v1
    buf[:b, d:2 * d] = v2
    yy = torch.full((rows,), -1, dtype=torch.int64, device=v1.device)
    yy[:b] = y
    buf[:, 2 * d:] = yy.view(-1, 1).view(torch.float32).view(rows, 2)
    return buf


def unpack_anchor_rows(buf: torch.Tensor, counts: list[int], rows: int, d: int):
    """Inverse of pack_anchor_rows over the gathered [world*rows, 2D+2] buffer -> (v1[B,D], v2[B,D], y[B])."""
    keep = torch.cat([torch.arange(r * rows, r * rows + c, device=buf.device) for r, c in enumerate(counts)])
    sel = buf.index_select(0, keep)
    y = sel[:, 2 * d:].contiguous().view(torch.int64).view(-1)
    return sel[:, :d].contiguous(), sel[:, d:2 * d].contiguous(), y


class _GatherAnchors(torch.autograd.Function):
    """All-gather of the local (v1, v2) rows; backward hands each rank the gradient rows of its own anchors
    (the gradients are already summed over ranks by the packed all-reduce of the step)."""

    @staticmethod
    def forward(ctx, v1, v2, y, owner):
        ctx.owner = owner
        ctx.b_loc = v1.shape[0]
        g1, g2, gy = owner._gather(v1, v2, y)
        ctx.mark_non_differentiable(gy)
        return g1, g2, gy

    @staticmethod
    def backward(ctx, d1, d2, _dy):
        lo = ctx.owner._anchor_offset
        return d1[lo:lo + ctx.b_loc].contiguous(), d2[lo:lo + ctx.b_loc].contiguous(), None, None


class PeerExchange:
    """NVLink peer-memory exchange buffers for one group of ranks on one box (CUDA IPC; ``crdpn_p2p_*``).

    Setup is collective: every rank allocates its buffer, the 64-byte IPC handles travel through
    ``dist.all_gather_object``, every rank maps its peers' buffers.  ``allgather`` / ``allreduce`` then are single
    kernel launches on the current stream (no NCCL call, CUDA-graph capturable)."""

    def __init__(self, group, rank, world, device, Bmax, Dmax):
        lib = _native.lib()
        self.rank, self.world, self.device, self.Bmax, self.Dmax = rank, world, device, int(Bmax), int(Dmax)
        n = ctypes.c_size_t(0)
        _native.check(lib.crdpn_p2p_buffer_bytes(self.Bmax, self.Dmax, world, ctypes.byref(n)), "crdpn_p2p_buffer_bytes")
        own = ctypes.c_void_p()
        with _native.on_device(device):
            _native.check(lib.crdpn_p2p_alloc(n.value, ctypes.byref(own)), "crdpn_p2p_alloc")
            handle = ctypes.create_string_buffer(64)
            _native.check(lib.crdpn_p2p_export(own, handle), "crdpn_p2p_export")
            handles = [handle.raw]
            if world > 1:
                handles = [None] * world
                dist.all_gather_object(handles, handle.raw, group=group)
            self._own = own
            self._imported = []
            ptrs = []
            for r in range(world):
                if r == rank:
                    ptrs.append(own.value)
                    continue
                peer = ctypes.c_void_p()
                _native.check(lib.crdpn_p2p_import(ctypes.create_string_buffer(handles[r], 64), ctypes.byref(peer)),
                              "crdpn_p2p_import")
                self._imported.append(peer)
                ptrs.append(peer.value)
        self._ptrs = (ctypes.c_void_p * world)(*ptrs)
        if world > 1:
            dist.barrier(group=group)  # nobody starts writing before every mapping exists

    def allgather(self, v1, v2, y, counts):
        B, D = sum(counts), v1.shape[1]
        offs = [0]
        for c in counts:
            offs.append(offs[-1] + c)
        out1 = torch.empty(B, D, dtype=torch.float32, device=v1.device)
        out2 = torch.empty(B, D, dtype=torch.float32, device=v1.device)
        outy = torch.empty(B, dtype=torch.int64, device=v1.device)
        offs_c = (ctypes.c_int32 * (self.world + 1))(*offs)
        with _native.on_device(v1.device):
            rc = _native.lib().crdpn_p2p_allgather_anchors(
                v1.data_ptr(), v2.data_ptr(), y.data_ptr(), D, offs_c, self._ptrs, self.rank, self.world,
                self.Bmax, self.Dmax, out1.data_ptr(), out2.data_ptr(), outy.data_ptr(), _stream_ptr(v1.device))
        _native.check(rc, "crdpn_p2p_allgather_anchors")
        return out1, out2, outy

    def allreduce(self, partial, tail_f64=None):
        """Sum of `partial` (fp32) over the ranks, in rank order; `tail_f64` (optional fp64 vector) rides behind it
        as fp32 words, so the step's result scalars need no separate cast / copy kernel."""
        n_tail = 0 if tail_f64 is None else tail_f64.numel()
        out = torch.empty(partial.numel() + n_tail, dtype=torch.float32, device=partial.device)
        with _native.on_device(partial.device):
            rc = _native.lib().crdpn_p2p_allreduce_f32(partial.data_ptr(), partial.numel(),
                                                       tail_f64.data_ptr() if n_tail else None, n_tail,
                                                       out.data_ptr(), self._ptrs,
                                                       self.rank, self.world, self.Bmax, self.Dmax,
                                                       _stream_ptr(partial.device))
        _native.check(rc, "crdpn_p2p_allreduce_f32")
        return out

    def close(self):
        lib = _native.lib()
        if getattr(self, "_own", None) is None:
            return
        torch.cuda.synchronize(self.device)
        with _native.on_device(self.device):
            for peer in self._imported:
                lib.crdpn_p2p_close(peer)
            lib.crdpn_p2p_free(self._own)
        self._own, self._imported = None, []


class ShardedContrastMemory(ContrastMemory):
    """ContrastMemory holding rows [row_begin,row_end) of the global banks; collectives on ``group``."""

    def __init__(self, inputSize, outputSize, K, T=0.07, momentum=0.5, group=None, rank=None, world_size=None,
                 local_negatives=False, comm="dist", fixed_local_batch=False, **kw):
        """comm: "dist" = torch.distributed collectives (NCCL on GPUs, gloo in the CPU tests);
        "p2p" = the per-step exchanges as kernels over NVLink peer memory (``PeerExchange``; the ranks must then stay
        within CRDPN_P2P_TIMEOUT_S seconds of each other, default 600 -- see include/crdpn_b200.h).
        fixed_local_batch: the per-rank batch sizes are exchanged ONCE (first call) instead of every call; a rank whose
        batch size then changes raises instead of hanging the next collective.  Leave it False when the last batch of
        an epoch may be ragged."""
        if comm not in ("dist", "p2p"):
            raise ValueError("comm must be 'dist' or 'p2p'")
        self.comm = comm
        self._px = None
        self.group = group
        self.world_size = dist.get_world_size(group) if world_size is None else world_size
        self.rank = dist.get_rank(group) if rank is None else rank
        self.fixed_local_batch = bool(fixed_local_batch)
        lo, hi = shard_bounds(outputSize, self.world_size, self.rank)
        if kw.get("seed") is None and self.world_size > 1 and dist.is_available() and dist.is_initialized():
            # replicated negatives (idx=None, local_negatives=False) must be the SAME list on every rank: take rank 0's
            # sampler seed everywhere (per-process torch.initial_seed() differs when a script seeds by rank)
            box = [int(torch.initial_seed()) if self.rank == 0 else None]
            dist.broadcast_object_list(box, src=dist.get_global_rank(group, 0) if group is not None else 0, group=group)
            kw["seed"] = box[0]
        super().__init__(inputSize, outputSize, K, T, momentum, row_begin=lo, row_end=hi, **kw)
        self.local_negatives = local_negatives
        if local_negatives:
            self.k_total = K * self.world_size  # every rank contributes K in-shard negatives per anchor
        self._counts = None

    # -- collectives (the only places that talk to torch.distributed) ---------------------------------------
    def _all_reduce(self, t):
        if self.world_size > 1:
            dist.all_reduce(t, op=dist.ReduceOp.SUM, group=self.group)
        return t

    def _all_gather_rows(self, buf):
        out = torch.empty(self.world_size * buf.shape[0], buf.shape[1], dtype=buf.dtype, device=buf.device)
        if self.world_size > 1:
            dist.all_gather_into_tensor(out, buf, group=self.group)
        else:
            out.copy_(buf)
        return out

    def _reduce_sums(self, res):
        return self._all_reduce(res)

    def _grad_buffers(self, v1, v2):
        # one fp32 buffer [grad_v1 | grad_v2 | 8 scalars]: the kernel writes the gradients straight into it
        B, D = v1.shape
        self._packed = torch.empty(2 * B * D + 8, dtype=torch.float32, device=v1.device)
        return self._packed[:B * D].view(B, D), self._packed[B * D:2 * B * D].view(B, D)

    def _reduce_partials(self, res, g1, g2):
        B, D = g1.shape
        packed = getattr(self, "_packed", None)
        if packed is None or g1.data_ptr() != packed.data_ptr():  # gradients were not produced in place
            packed = torch.empty(2 * B * D + 8, dtype=torch.float32, device=g1.device)
            packed[:B * D] = g1.reshape(-1)
            packed[B * D:2 * B * D] = g2.reshape(-1)
        if self.comm == "p2p":
            # one kernel over NVLink peer memory; the 8 fp64 result scalars ride along as fp32 words
            packed = self._peer_exchange(B, D, g1.device).allreduce(packed[:2 * B * D], res)
        else:
            packed[2 * B * D:] = res  # fp64 -> fp32 (1e-7 relative on the loss partials)
            self._all_reduce(packed)  # ONE packed exchange after the kernel
        return packed[2 * B * D:], packed[:B * D].view(B, D), packed[B * D:2 * B * D].view(B, D)

    def _peer_exchange(self, B, D, device):
        if self._px is None or self._px.Bmax < B or self._px.Dmax < D:
            if self._px is not None:
                self._px.close()
            self._px = PeerExchange(self.group, self.rank, self.world_size, device, max(B, 64), max(D, 128))
        return self._px

    def _ensure_counts(self, b_loc, device):
        """Per-rank batch sizes.  The exchange is a collective, so it has to be issued by EVERY rank or by none: it runs
        on every call (one tiny all-reduce + a host read) unless ``fixed_local_batch`` promises constant sizes, in
        which case it runs once and a later change raises on the rank that sees it."""
        if self.fixed_local_batch and self._counts is not None:
            if self._counts[self.rank] != b_loc:
                raise RuntimeError(f"fixed_local_batch=True but the local batch changed from {self._counts[self.rank]} to "
                                   f"{b_loc}; call reset_batch_sizes() on every rank first")
            return self._counts
        cnt = torch.zeros(self.world_size, dtype=torch.int64, device=device)
        cnt[self.rank] = b_loc
        self._counts = self._all_reduce(cnt).tolist()
        return self._counts

    def reset_batch_sizes(self):
        """Forget the cached per-rank batch sizes (collective in effect: call it on every rank)."""
        self._counts = None

    def _ensure_local_sampler(self, device):
        """In-shard negatives come from this rank's own Philox stream."""
        if getattr(self, "_local_sampler", None) is None:
            from .crd import AliasMethod
            self._local_sampler = AliasMethod(torch.ones(self.row_end - self.row_begin), seed=self.multinomial.seed + 7919 * (self.rank + 1))
            self._local_sampler.to(device)
        return self._local_sampler

    def _sampler_for_offset(self):
        return getattr(self, "_local_sampler", None) or self.multinomial

    NO_COMPACT = 0x20   # variant bit: no filter pre-pass (every entry of the list lives in this shard anyway)

    def _step_variant(self, B, K1, D):
        v = super()._step_variant(B, K1, D)
        if self.local_negatives and not (v & self.STREAM) and not (self.variant & 0x40):
            v |= self.NO_COMPACT   # in-shard negatives: nothing to filter out (with SWEEP: band sort without survivor compaction)
        return v

    def _gather(self, v1, v2, y):
        """ONE packed exchange before the kernel: local anchors -> all anchors (uneven B_loc allowed)."""
        counts = self._ensure_counts(v1.shape[0], v1.device)
        rows = max(counts)
        self._anchor_offset = sum(counts[:self.rank])
        d = v1.shape[1]
        if self.comm == "p2p":
            return self._peer_exchange(sum(counts), d, v1.device).allgather(
                v1.contiguous(), v2.contiguous(), y.contiguous().to(torch.int64), counts)
        if min(counts) == rows:  # even split: no padding, no compaction
            buf = torch.cat([v1, v2, y.contiguous().view(-1, 1).view(torch.float32)], dim=1)
            out = self._all_gather_rows(buf)
            return (out[:, :d].contiguous(), out[:, d:2 * d].contiguous(),
                    out[:, 2 * d:].contiguous().view(torch.int64).view(-1))
        gathered = self._all_gather_rows(pack_anchor_rows(v1, v2, y, rows))
        return unpack_anchor_rows(gathered, counts, rows, d)

    def _prepare(self, v1, v2, y, idx):
        if idx is None and self.local_negatives:
            # in-shard negatives from this rank's own Philox stream; column 0 stays the global positive index
            K1 = self._host_params().K + 1
            idx = self._ensure_local_sampler(v1.device).draw_contrast(y.contiguous().to(torch.int64), K1, row_base=self.row_begin)
        return super()._prepare(v1, v2, y, idx)

    def step_resident(self, v1_loc, v2_loc, y_loc, contrast_idx, out=None):
        """The sharded step on device-resident LOCAL embeddings through ``crdpn_crd_step_sharded`` (peer-memory
        all-gather -> scoring pass over this shard -> reduction + momentum update + sum over ranks in one kernel:
        3 launches, no NCCL call, CUDA-graph capturable).  Z must be frozen.  Returns ``out``: dict with ``reduced``
        [2*B*D + 8] f32 (grad_v1 | grad_v2 | result words, word 5 = loss of the whole batch), ``v1_all``, ``v2_all``,
        ``y_all``; pass the returned dict back in to reuse its buffers (needed for graph capture)."""
        if self.comm != "p2p":
            raise RuntimeError("step_resident needs comm='p2p'")
        hp = self._host_params()
        if hp.Z1 <= 0 or hp.Z2 <= 0:
            raise RuntimeError("step_resident: freeze Z first (one ordinary forward)")
        from .crd import EPS
        dev = v1_loc.device
        D = v1_loc.shape[1]
        counts = self._ensure_counts(v1_loc.shape[0], dev)
        B = sum(counts)
        K1 = contrast_idx.shape[1]
        px = self._peer_exchange(B, D, dev)
        if out is None:
            out = dict(reduced=torch.empty(2 * B * D + 8, dtype=torch.float32, device=dev),
                       partial=torch.empty(2 * B * D, dtype=torch.float32, device=dev),
                       result=torch.empty(8, dtype=torch.float64, device=dev),
                       v1_all=torch.empty(B, D, dtype=torch.float32, device=dev),
                       v2_all=torch.empty(B, D, dtype=torch.float32, device=dev),
                       y_all=torch.empty(B, dtype=torch.int64, device=dev))
            offs = [0]
            for c in counts:
                offs.append(offs[-1] + c)
            out["offs"] = (ctypes.c_int32 * (self.world_size + 1))(*offs)
        m1, m2, stride, dt = self._banks()
        variant = self._step_variant(B, K1, D)
        ws = self._workspace(B, K1, D, dev, variant)
        if contrast_idx.dtype == torch.int32 and not (variant & self.STREAM):
            variant |= self.IDX32
        elif contrast_idx.dtype != torch.int64:
            raise RuntimeError("step_resident: contrast_idx must be int64 (or int32 for the gather kernels)")
        with _native.on_device(dev):
            rc = _native.lib().crdpn_crd_step_sharded(
                m1.data_ptr(), m2.data_ptr(), stride, dt, v1_loc.data_ptr(), v2_loc.data_ptr(), y_loc.data_ptr(),
                out["offs"], px._ptrs, self.rank, self.world_size, px.Bmax, px.Dmax, contrast_idx.data_ptr(),
                K1, D, self.nLem, self.k_total, self.row_begin, self.row_end, hp.T, hp.Z1, hp.Z2, EPS, hp.m, 1.0 - hp.m,
                out["v1_all"].data_ptr(), out["v2_all"].data_ptr(), out["y_all"].data_ptr(), out["partial"].data_ptr(),
                out["result"].data_ptr(), out["reduced"].data_ptr(), ws.data_ptr(), ws.numel(), variant, _stream_ptr(dev))
        _native.check(rc, "crdpn_crd_step_sharded")
        return out

    def fused_loss(self, v1, v2, y, idx=None):
        """v1, v2, y: this rank's LOCAL anchors; idx: replicated [B, K+1] or per-rank [B, K_loc+1]."""
        g1, g2, gy = _GatherAnchors.apply(v1, v2, y, self)
        g1c, g2c, gy, idx = self._prepare(g1, g2, gy, idx)
        return _FusedCRDFunction.apply(g1c, g2c, gy, idx, self)


class _ShardedCRDLossFunction(torch.autograd.Function):
    """The sharded step with the peer-memory exchanges as one autograd node and two foreign calls
    (``crdpn_crd_loss_forward_sharded``: 6 launches, the sum over ranks fused into the reduction kernel;
    ``crdpn_crd_loss_backward`` on the local rows: 2 launches)."""

    @staticmethod
    def forward(ctx, f_s, f_t, Ws, bs, Wt, bt, y, contrast_idx, crit):
        from .crd import EPS
        mem = crit.contrast
        dev = f_s.device
        xs, xt = f_s.detach(), f_t.detach()
        if xs.dim() != 2 or not xs.is_contiguous():
            xs = xs.reshape(xs.shape[0], -1).contiguous()
        if xt.dim() != 2 or not xt.is_contiguous():
            xt = xt.reshape(xt.shape[0], -1).contiguous()
        Wsc, bsc, Wtc, btc = (t.detach() if t.is_contiguous() else t.detach().contiguous() for t in (Ws, bs, Wt, bt))
        B_loc, D = xs.shape[0], Wsc.shape[0]
        yc = y if (y.dtype == torch.int64 and y.is_contiguous()) else y.contiguous().to(torch.int64)
        counts = mem._ensure_counts(B_loc, dev)
        B = sum(counts)
        a0 = sum(counts[:mem.rank])
        hp = mem._host_params()
        K1 = hp.K + 1
        px = mem._peer_exchange(B, D, dev)
        offs = [0]
        for c in counts:
            offs.append(offs[-1] + c)
        offs_c = (ctypes.c_int32 * (mem.world_size + 1))(*offs)
        BD, BlD = B * D, B_loc * D
        Bp = (B_loc + 3) & ~3
        # one fp32 arena: [result(16) | pre_s | pre_t | v1_loc | v2_loc | inv1 | inv2 | v1_all | v2_all | partial(2BD) | reduced(2BD+8)]
        # (carried to backward as a plain attribute: the returned loss is one of its words, and an in-place op on the
        # loss must not trip autograd's version check on the gradient rows next to it)
        arena = torch.empty(16 + 4 * BlD + 2 * Bp + 2 * BD + 2 * BD + 2 * BD + 8, dtype=torch.float32, device=dev)
        base = arena.data_ptr()
        o_pre_s, o_pre_t, o_v1l, o_v2l = (base + 4 * (16 + i * BlD) for i in range(4))
        o_inv1 = base + 4 * (16 + 4 * BlD)
        o_inv2 = o_inv1 + 4 * Bp
        f_all = 16 + 4 * BlD + 2 * Bp
        o_v1a, o_v2a, o_part = base + 4 * f_all, base + 4 * (f_all + BD), base + 4 * (f_all + 2 * BD)
        f_red = f_all + 4 * BD
        o_red = base + 4 * f_red
        y_all = torch.empty(B, dtype=torch.int64, device=dev)
        dev_offset = None
        m1, m2, stride, dt = mem._banks()
        variant = mem._step_variant(B, K1, D)
        ws = mem._workspace(B, K1, D, dev, variant)
        if contrast_idx is not None:
            mem._check_device(contrast_idx, "contrast_idx")
            if contrast_idx.dtype == torch.int32 and not (variant & mem.STREAM):
                variant |= mem.IDX32                      # consumed as it is: half the bytes to copy and scan
            else:
                contrast_idx = contrast_idx.to(torch.int64)
            contrast_idx = contrast_idx.contiguous()
            if contrast_idx.shape != (B, K1):
                raise RuntimeError(f"contrast_idx must have shape [B, K+1] = {(B, K1)}, got {tuple(contrast_idx.shape)}")
            cidx_ptr, scratch_ptr, tables, seed, offset = contrast_idx.data_ptr(), None, (None, None), 0, 0
        else:
            smp = mem._ensure_local_sampler(dev)
            if smp.uniform and not (variant & mem.STREAM):
                scratch_ptr = None                        # the scoring pass draws the in-shard negatives itself
                dev_offset = mem._device_offset(dev)
                if dev_offset is not None:                # CUDA-graph replays: advancing part of the offset on the device
                    variant |= mem.DEVICE_OFFSET
                    scratch_ptr = dev_offset.data_ptr()
            else:
                scratch = mem._idx_scratch
                if scratch is None or scratch.numel() != B * K1 or scratch.device != dev:
                    scratch = mem._idx_scratch = torch.empty(B * K1, dtype=torch.int64, device=dev)
                scratch_ptr = scratch.data_ptr()
            cidx_ptr, tables, seed, offset = None, smp.table_ptrs(), smp.seed, smp.offset
        with _native.on_device(dev):
            rc = _native.lib().crdpn_crd_loss_forward_sharded(
                xs.data_ptr(), xs.shape[1], Wsc.data_ptr(), bsc.data_ptr(), xt.data_ptr(), xt.shape[1], Wtc.data_ptr(), btc.data_ptr(),
                yc.data_ptr(), offs_c, px._ptrs, mem.rank, mem.world_size, px.Bmax, px.Dmax,
                cidx_ptr, tables[0], tables[1], seed, offset, scratch_ptr,
                m1.data_ptr(), m2.data_ptr(), stride, dt, K1, D, mem.nLem, mem.k_total, mem.row_begin, mem.row_end,
                hp.T, hp.Z1, hp.Z2, EPS, hp.m, 1.0 - hp.m,
                o_pre_s, o_pre_t, o_v1l, o_v2l, o_inv1, o_inv2, o_v1a, o_v2a, y_all.data_ptr(), o_part, base, o_red,
                ws.data_ptr(), ws.numel(), variant, _stream_ptr(dev))
        _native.check(rc, "crdpn_crd_loss_forward_sharded")
        if contrast_idx is None and dev_offset is None:
            smp.offset += B * K1
        ctx.save_for_backward(xs, xt, Wsc, Wtc)
        ctx.arena = arena
        ctx.geom = (B, B_loc, D, a0, f_red, f_s.shape, f_t.shape)
        ctx.need_dx = (ctx.needs_input_grad[0], ctx.needs_input_grad[1])
        return arena[f_red + 2 * BD + 5]   # loss_s + loss_t of the WHOLE batch, summed over ranks

    @staticmethod
    def backward(ctx, grad_out):
        xs, xt, Ws, Wt = ctx.saved_tensors
        arena = ctx.arena
        B, B_loc, D, a0, f_red, shp_s, shp_t = ctx.geom
        dev = xs.device
        BD, BlD = B * D, B_loc * D
        Bp = (B_loc + 3) & ~3
        base = arena.data_ptr()
        o_v1l, o_v2l = base + 4 * (16 + 2 * BlD), base + 4 * (16 + 3 * BlD)
        o_inv1 = base + 4 * (16 + 4 * BlD)
        o_inv2 = o_inv1 + 4 * Bp
        o_g1 = base + 4 * (f_red + a0 * D)            # this rank's rows of the rank-summed gradients
        o_g2 = base + 4 * (f_red + BD + a0 * D)
        scale = grad_out.detach().to(torch.float32).contiguous()
        dWs, dWt = torch.empty_like(Ws), torch.empty_like(Wt)
        dbs = torch.empty(2 * D, dtype=torch.float32, device=dev)
        dxs = torch.empty_like(xs) if ctx.need_dx[0] else None
        dxt = torch.empty_like(xt) if ctx.need_dx[1] else None
        d_pre = torch.empty(2 * BlD, dtype=torch.float32, device=dev)
        with _native.on_device(dev):
            rc = _native.lib().crdpn_crd_loss_backward(
                xs.data_ptr(), xs.shape[1], Ws.data_ptr(), o_v1l, o_inv1, o_g1,
                xt.data_ptr(), xt.shape[1], Wt.data_ptr(), o_v2l, o_inv2, o_g2,
                scale.data_ptr(), B_loc, D,
                dWs.data_ptr(), dbs.data_ptr(), dxs.data_ptr() if dxs is not None else None,
                dWt.data_ptr(), dbs.data_ptr() + 4 * D, dxt.data_ptr() if dxt is not None else None,
                d_pre.data_ptr(), _stream_ptr(dev))
        _native.check(rc, "crdpn_crd_loss_backward")
        return (dxs.view(shp_s) if dxs is not None else None, dxt.view(shp_t) if dxt is not None else None,
                dWs, dbs[:D], dWt, dbs[D:], None, None, None)


class ShardedCRDLoss(nn.Module):
    """CRDLoss over row-sharded banks.  forward(f_s_local, f_t_local, idx_local, contrast_idx) -> global loss.

    The returned loss is the loss of the WHOLE batch (identical on every rank); its gradient w.r.t. the local
    features and this rank's embed parameters covers the local anchors only, so embed-parameter gradients must be
    SUMMED over ranks (``allreduce_embed_grads``) to equal the single-GPU gradients."""

    def __init__(self, opt, group=None, rank=None, world_size=None, local_negatives=False, comm="dist",
                 fixed_local_batch=False, **memory_kwargs):
        super().__init__()
        self.embed_s = Embed(opt.s_dim, opt.feat_dim)
        self.embed_t = Embed(opt.t_dim, opt.feat_dim)
        self.contrast = ShardedContrastMemory(opt.feat_dim, opt.n_data, opt.nce_k, opt.nce_t, opt.nce_m, group=group,
                                              rank=rank, world_size=world_size, local_negatives=local_negatives,
                                              comm=comm, fixed_local_batch=fixed_local_batch, **memory_kwargs)

    def forward(self, f_s, f_t, idx, contrast_idx=None):
        mem = self.contrast
        hp = mem._host_params()
        fast = (mem.comm == "p2p" and hp.Z1 > 0 and hp.Z2 > 0 and f_s.is_cuda and f_t.is_cuda
                and f_s.dtype == torch.float32 and f_t.dtype == torch.float32
                and (contrast_idx is not None or mem.local_negatives))
        if fast:   # (the first call freezes Z through the general path below)
            return _ShardedCRDLossFunction.apply(f_s, f_t, self.embed_s.linear.weight, self.embed_s.linear.bias,
                                                 self.embed_t.linear.weight, self.embed_t.linear.bias, idx, contrast_idx, self)
        return mem.fused_loss(self.embed_s(f_s), self.embed_t(f_t), idx, contrast_idx)

    def allreduce_embed_grads(self):
        grads = [p.grad for p in list(self.embed_s.parameters()) + list(self.embed_t.parameters()) if p.grad is not None]
        if grads and self.contrast.world_size > 1:
            flat = torch.cat([g.reshape(-1) for g in grads])
            dist.all_reduce(flat, group=self.contrast.group)
            off = 0
            for g in grads:
                g.copy_(flat[off:off + g.numel()].view_as(g))
                off += g.numel()
