"""Tail of ``PoseEstimator`` -- everything between the two encoders and the losses -- as ONE kernel launch.

Reference: ``auxiliary/model.py:183-203`` (``DeformNet``: four 1x1 Conv1d on a length-1 sequence = four Linear layers, three
BatchNorm1d + ReLU, tanh) and ``model.py:238-272`` (``cat`` of the shape and image features, the six ``fc_*`` heads, the
``projector`` MLP); eval-mode call site ``KD/common/base_class.py:363`` (frozen teacher, ``base_class.py:317``), train-mode
call site ``training.py:47`` (``model.train()`` at ``training.py:30``, backward at ``training.py:75``).

Two modules over the same kernel (``csrc/pose_tail.cu``: a persistent tcgen05 chain -- every layer a set of
(128-output tile, K-split) tasks over all SMs, layers handing over through device-side counters, weights streamed once as
bf16 (hi, lo) operand images, three MMAs per product = fp32-accurate):

* ``FrozenPoseTail`` -- the KD loop's frozen teacher: BatchNorm (running statistics) folded into the packed weights, the
  concat as the first layer's K range, the six heads as one 120-wide layer.  One launch per call, no gradient path.
* ``PoseTail`` -- the trainable tail with the reference's own parameter names (``deformNet.conv1.weight`` ...
  ``projector.6.bias``), so a ``PoseEstimator`` checkpoint loads by key.  ``.eval()`` runs the folded chain; ``.train()``
  runs the same kernel with batch-statistics BatchNorm in its reduce step (running statistics updated in place) and
  differentiates it: the backward is one more foreign call (``crdpn_pose_tail_backward``: per layer the activation /
  BatchNorm pull-back, ``dW = g^T x`` and ``dx = g W`` as fp32 FFMA kernels on the saved activations).

There is no CPU path: non-CUDA inputs raise.
"""
from __future__ import annotations

import ctypes

import torch
from torch import nn

from . import _native

HEADS = ("fc_cls_azi", "fc_cls_ele", "fc_cls_inp", "fc_reg_azi", "fc_reg_ele", "fc_reg_inp")
BF16_FLAG, TRAIN_FLAG = 1, 2
NONE, RELU, TANH = 0, 1, 2


_Layer = _native.PoseTailLayer


def _bind():
    return _native.lib()


def _aligned(nbytes: int, device, zero=False) -> torch.Tensor:
    """A uint8 view of `nbytes` whose address is a multiple of 1024 (operand images / the workspace need that)."""
    raw = (torch.zeros if zero else torch.empty)(nbytes + 1024, dtype=torch.uint8, device=device)
    off = (-raw.data_ptr()) % 1024
    return raw[off:off + nbytes]


def _ptr(t):
    return None if t is None else t.data_ptr()


class _Chain:
    """Packed weight images + layer table of one chain on one device; runs ``crdpn_pose_tail_forward``."""

    def __init__(self, spec, shape_dim: int, img_dim: int, device):
        # spec: list of dicts(O, I, src, act)
        self.spec, self.shape_dim, self.img_dim, self.device = spec, shape_dim, img_dim, device
        lib = _bind()
        self.images = []
        for s in spec:
            n = ctypes.c_size_t()
            _native.check(lib.crdpn_pose_tail_image_bytes(s["O"], s["I"], ctypes.byref(n)), "crdpn_pose_tail_image_bytes")
            self.images.append(_aligned(n.value, device))
        self._ws = {}

    def pack(self, l: int, W: torch.Tensor, row_scale=None):
        W = W.detach()
        if W.dim() == 3:
            W = W[:, :, 0]
        W = W.to(torch.float32).contiguous()
        s = self.spec[l]
        if tuple(W.shape) != (s["O"], s["I"]):
            raise RuntimeError(f"pose tail layer {l}: weight shape {tuple(W.shape)} != {(s['O'], s['I'])}")
        rs = None if row_scale is None else row_scale.detach().to(torch.float32).contiguous()
        with _native.on_device(self.device):
            rc = _bind().crdpn_pose_tail_pack_weights(W.data_ptr(), _ptr(rs), s["O"], s["I"], self.images[l].data_ptr(),
                                                      _native.stream_ptr(self.device))
        _native.check(rc, "crdpn_pose_tail_pack_weights")

    def table(self, biases, outs, bn=None):
        arr = (_Layer * len(self.spec))()
        for l, s in enumerate(self.spec):
            a = arr[l]
            a.weights, a.bias, a.O, a.I, a.src, a.act = self.images[l].data_ptr(), biases[l].data_ptr(), s["O"], s["I"], s["src"], s["act"]
            a.out = _ptr(outs[l])
            if bn is not None and bn[l] is not None:
                g = bn[l]
                a.gamma, a.beta, a.running_mean, a.running_var = _ptr(g["gamma"]), _ptr(g["beta"]), _ptr(g["rm"]), _ptr(g["rv"])
                a.save_mean, a.save_istd, a.xhat = _ptr(g["mean"]), _ptr(g["istd"]), _ptr(g["xhat"])
        return arr

    def workspace(self, arr, B: int):
        ws = self._ws.get(B)
        if ws is None:
            n = ctypes.c_size_t()
            with _native.on_device(self.device):
                _native.check(_bind().crdpn_pose_tail_workspace_bytes(arr, len(self.spec), B, self.shape_dim, self.img_dim,
                                                                      ctypes.byref(n)), "crdpn_pose_tail_workspace_bytes")
            ws = self._ws[B] = _aligned(n.value, self.device, zero=True)   # the hand-over counters start at zero
        return ws

    def run(self, arr, sf, img, flags=0, momentum=0.1, eps=1e-5):
        B = img.shape[0]
        ws = self.workspace(arr, B)
        with _native.on_device(self.device):
            rc = _bind().crdpn_pose_tail_forward(arr, len(self.spec), _ptr(sf), img.data_ptr(), B, self.shape_dim, self.img_dim,
                                                 flags, momentum, eps, ws.data_ptr(), ws.numel(), _native.stream_ptr(self.device))
        _native.check(rc, "crdpn_pose_tail_forward")


def _tail_spec(shape_dim, img_dim, widths, head_total, proj):
    """Layer table of PoseEstimator's tail: DeformNet (4 layers), the concatenated heads, the projector (3 layers)."""
    C = shape_dim + img_dim
    w1, w2, w3, w4 = widths
    p1, p2, p3 = proj
    return [dict(O=w1, I=C, src=-1, act=RELU), dict(O=w2, I=w1, src=0, act=RELU), dict(O=w3, I=w2, src=1, act=RELU),
            dict(O=w4, I=w3, src=2, act=TANH), dict(O=head_total, I=w4, src=3, act=NONE),
            dict(O=p1, I=img_dim, src=-2, act=RELU), dict(O=p2, I=p1, src=5, act=RELU), dict(O=p3, I=p2, src=6, act=NONE)]


_BN_OF = {0: "deformNet.bn1", 1: "deformNet.bn2", 2: "deformNet.bn3", 5: "projector.1", 6: "projector.4"}
_LIN_OF = {0: "deformNet.conv1", 1: "deformNet.conv2", 2: "deformNet.conv3", 3: "deformNet.conv4", 5: "projector.0",
           6: "projector.3", 7: "projector.6"}


def _check_inputs(shape_feature, img_feature, shape_dim, img_dim, who):
    if not (shape_feature.is_cuda and img_feature.is_cuda):
        raise RuntimeError(f"{who} inputs must be CUDA tensors: this package has no CPU fallback")
    if shape_feature.dim() != 2 or img_feature.dim() != 2 or shape_feature.shape[1] != shape_dim or img_feature.shape[1] != img_dim \
            or shape_feature.shape[0] != img_feature.shape[0]:
        raise RuntimeError(f"{who}: expected shape_feature [B, {shape_dim}] and img_feature [B, {img_dim}]")


class FrozenPoseTail(nn.Module):
    """forward(shape_feature [B, Fs], img_feature [B, Fi]) -> ([cls_azi, cls_ele, cls_inp, reg_azi, reg_ele, reg_inp], x [B, 200],
    projector(img_feature) [B, 200]) -- the three values ``PoseEstimator.forward`` returns (``model.py:272``), for a frozen
    ``.eval()`` teacher.  Build it from a trained ``PoseEstimator``'s ``state_dict`` (``from_state_dict``); call ``refold()``
    after loading new weights.  ``dtype=torch.float32`` (default): fp32-accurate split arithmetic, outputs within 1e-5 of the
    reference's; ``torch.bfloat16``: one bf16 MMA per product on half the weight bytes (north_star's 1e-2 tolerance mode).
    The returned tensors are reused by the next call with the same batch size.  ``max_rows``: larger batches run in chunks."""

    MAX_ROWS = 256

    def __init__(self, sd: dict, prefix: str = "", dtype=torch.float32, eps: float = 1e-5, graph: bool = True):
        super().__init__()
        if dtype not in (torch.float32, torch.bfloat16):
            raise ValueError("FrozenPoseTail dtype must be float32 (reference parity) or bfloat16 (1e-2 tolerance mode)")
        self.dtype, self.eps = dtype, eps
        self._source = (sd, prefix)
        g = lambda k: sd[prefix + k].detach().to(torch.float32)
        self.head_sizes = [int(sd[prefix + h + ".weight"].shape[0]) for h in HEADS]
        self.img_dim = int(g("projector.0.weight").shape[1])
        self.shape_dim = int(g("deformNet.conv1.weight").shape[1]) - self.img_dim
        widths = [int(g(f"deformNet.conv{n}.weight").shape[0]) for n in (1, 2, 3, 4)]
        proj = [int(g(f"projector.{n}.weight").shape[0]) for n in (0, 3, 6)]
        self.spec = _tail_spec(self.shape_dim, self.img_dim, widths, sum(self.head_sizes), proj)
        for l, s in enumerate(self.spec):   # folded fp32 weights / biases live as buffers so .to(device) moves them
            W, b, scale = self._folded(l)
            self.register_buffer(f"W{l}", W, persistent=False)
            self.register_buffer(f"b{l}", b, persistent=False)
            self.register_buffer(f"s{l}", scale, persistent=False)
        self._chain = None
        self._outs = {}

    # -- construction ---------------------------------------------------------------------------------------
    def _folded(self, l):
        sd, prefix = self._source
        g = lambda k: sd[prefix + k].detach().to(torch.float32)
        if l == 4:
            W = torch.cat([g(h + ".weight") for h in HEADS], 0)
            b = torch.cat([g(h + ".bias") for h in HEADS], 0)
        else:
            W, b = g(_LIN_OF[l] + ".weight"), g(_LIN_OF[l] + ".bias")
            if W.dim() == 3:
                W = W[:, :, 0]
        scale = torch.ones(W.shape[0])
        if l in _BN_OF:
            bn = _BN_OF[l]
            scale = g(bn + ".weight") / torch.sqrt(g(bn + ".running_var") + self.eps)
            b = (b - g(bn + ".running_mean")) * scale + g(bn + ".bias")
        return W.contiguous(), b.contiguous(), scale.contiguous()

    @classmethod
    def from_state_dict(cls, sd: dict, prefix: str = "", graph: bool = True, dtype=torch.float32) -> "FrozenPoseTail":
        """``graph`` is accepted for compatibility with the library-GEMM version of this class (which needed a CUDA graph to
        be one launch); the chain kernel is a single launch by construction."""
        return cls(sd, prefix, dtype)

    def refold(self, sd: dict | None = None, prefix: str | None = None):
        """Re-derive the folded weights (after the teacher's weights changed) and repack the operand images."""
        if sd is not None:
            self._source = (sd, prefix or "")
        for l in range(len(self.spec)):
            W, b, scale = self._folded(l)
            getattr(self, f"W{l}").copy_(W); getattr(self, f"b{l}").copy_(b); getattr(self, f"s{l}").copy_(scale)
        self._chain = None
        return self

    def _apply(self, fn, *a, **k):
        self._chain, self._outs = None, {}
        return super()._apply(fn, *a, **k)

    def _ensure_chain(self, device):
        if self._chain is None or self._chain.device != device:
            ch = _Chain(self.spec, self.shape_dim, self.img_dim, device)
            for l in range(len(self.spec)):
                ch.pack(l, getattr(self, f"W{l}"), getattr(self, f"s{l}"))
            self._chain, self._outs = ch, {}
        return self._chain

    # -- forward --------------------------------------------------------------------------------------------
    def _run(self, sf, img):
        ch = self._ensure_chain(img.device)
        B = img.shape[0]
        ent = self._outs.get(B)
        if ent is None:
            outs = [None] * len(self.spec)
            for l in (3, 4, 7):
                outs[l] = torch.empty(B, self.spec[l]["O"], dtype=torch.float32, device=img.device)
            arr = ch.table([getattr(self, f"b{l}") for l in range(len(self.spec))], outs)
            ent = self._outs[B] = (arr, outs)
        arr, outs = ent
        ch.run(arr, sf, img, BF16_FLAG if self.dtype == torch.bfloat16 else 0)
        return outs[4], outs[3], outs[7]

    @torch.no_grad()
    def forward(self, shape_feature, img_feature):
        _check_inputs(shape_feature, img_feature, self.shape_dim, self.img_dim, "FrozenPoseTail")
        sf = shape_feature.detach().to(torch.float32).contiguous()
        img = img_feature.detach().to(torch.float32).contiguous()
        B = img.shape[0]
        if B <= self.MAX_ROWS:
            heads, x, p = self._run(sf, img)
        else:   # eval mode is row-independent: larger batches go through in chunks
            parts = [tuple(t.clone() for t in self._run(sf[i:i + self.MAX_ROWS].contiguous(), img[i:i + self.MAX_ROWS].contiguous()))
                     for i in range(0, B, self.MAX_ROWS)]
            heads, x, p = (torch.cat([q[j] for q in parts], 0) for j in range(3))
        return list(torch.split(heads, self.head_sizes, dim=1)), x, p


# =========================================================================================================================
class DeformNet(nn.Module):
    """Parameter container with the reference's names (``auxiliary/model.py:183-195``); the arithmetic runs in ``PoseTail``."""

    def __init__(self, bottleneck_size=1024):
        super().__init__()
        self.bottleneck_size = bottleneck_size
        self.conv1 = nn.Conv1d(bottleneck_size, bottleneck_size, 1)
        self.conv2 = nn.Conv1d(bottleneck_size, bottleneck_size // 2, 1)
        self.conv3 = nn.Conv1d(bottleneck_size // 2, bottleneck_size // 4, 1)
        self.conv4 = nn.Conv1d(bottleneck_size // 4, 200, 1)
        self.th = nn.Tanh()
        self.bn1 = nn.BatchNorm1d(bottleneck_size)
        self.bn2 = nn.BatchNorm1d(bottleneck_size // 2)
        self.bn3 = nn.BatchNorm1d(bottleneck_size // 4)


class _PoseTailTrainFunction(torch.autograd.Function):
    """Train-mode chain: forward = eight weight packs + one launch of the chain kernel with batch-statistics BatchNorm;
    backward = ``crdpn_pose_tail_backward`` (pull-backs, dW = g^T x, dx = g W on the activations the forward kept)."""

    @staticmethod
    def forward(ctx, tail, sf, img, *params):
        spec = tail.spec
        dev = img.device
        B = img.shape[0]
        ch = tail._ensure_chain(dev)
        lin, bns = tail._layer_params()
        for l in range(len(spec)):
            ch.pack(l, lin[l][0])
        outs = [torch.empty(B, s["O"], dtype=torch.float32, device=dev) for s in spec]
        bn = [None] * len(spec)
        for l, m in bns.items():
            O = spec[l]["O"]
            bn[l] = dict(gamma=m.weight.detach(), beta=m.bias.detach(),
                         rm=m.running_mean if m.track_running_stats else None, rv=m.running_var if m.track_running_stats else None,
                         mean=torch.empty(O, dtype=torch.float32, device=dev), istd=torch.empty(O, dtype=torch.float32, device=dev),
                         xhat=torch.empty(B, O, dtype=torch.float32, device=dev))
        biases = [lin[l][1].detach().contiguous() for l in range(len(spec))]
        arr = ch.table(biases, outs, bn)
        mom = next(iter(bns.values())).momentum if bns else 0.1
        ch.run(arr, sf, img, TRAIN_FLAG, 0.1 if mom is None else mom, tail.eps)
        for m in bns.values():
            if m.track_running_stats and m.num_batches_tracked is not None:
                m.num_batches_tracked += 1
        # every tensor the backward reads goes through save_for_backward: three of `outs` are this node's own OUTPUTS, and
        # an output kept as a plain ctx attribute closes a cycle node -> tensor -> grad_fn -> node that only a garbage-
        # collector pass breaks -- until then the step's activations stay allocated and the parameters' AccumulateGrad
        # nodes stay alive on the stream of that step (a later CUDA-graph capture of the same module then fails with
        # "would make the legacy stream depend on a capturing blocking stream")
        bn_layers = [l for l in range(len(spec)) if bn[l] is not None]
        ctx.tail, ctx.lin, ctx.bn_layers, ctx.n_layers = tail, lin, bn_layers, len(spec)
        ctx.save_for_backward(sf, img, *outs, *[bn[l][k] for l in bn_layers for k in ("xhat", "gamma", "istd")])
        return outs[4], outs[3], outs[7]

    @staticmethod
    def backward(ctx, g_heads, g_x, g_p):
        tail = ctx.tail
        saved = ctx.saved_tensors
        sf, img = saved[0], saved[1]
        outs = saved[2:2 + ctx.n_layers]
        bn = [None] * ctx.n_layers
        for i, l in enumerate(ctx.bn_layers):
            xhat, gamma, istd = saved[2 + ctx.n_layers + 3 * i:5 + ctx.n_layers + 3 * i]
            bn[l] = dict(xhat=xhat, gamma=gamma, istd=istd)
        spec = tail.spec
        dev = img.device
        B = img.shape[0]
        lin = ctx.lin
        f32 = lambda g: None if g is None else g.detach().to(torch.float32).contiguous()
        g_ext = {3: f32(g_x), 4: f32(g_heads), 7: f32(g_p)}
        arr = (_native.PoseTailBwdLayer * len(spec))()
        dW, db, dG, dB = {}, {}, {}, {}
        keep = []
        for l, s in enumerate(spec):
            W = lin[l][0].detach()
            if not W.is_contiguous():
                W = W.contiguous()
                keep.append(W)
            dW[l], db[l] = torch.empty_like(W), torch.empty(s["O"], dtype=torch.float32, device=dev)
            a = arr[l]
            a.W, a.y, a.O, a.I, a.src, a.act = W.data_ptr(), outs[l].data_ptr(), s["O"], s["I"], s["src"], s["act"]
            a.g_out = _ptr(g_ext.get(l))
            a.dW, a.db = dW[l].data_ptr(), db[l].data_ptr()
            if bn[l] is not None:
                dG[l], dB[l] = torch.empty_like(db[l]), torch.empty_like(db[l])
                a.xhat, a.gamma, a.istd = bn[l]["xhat"].data_ptr(), bn[l]["gamma"].data_ptr(), bn[l]["istd"].data_ptr()
                a.dgamma, a.dbeta = dG[l].data_ptr(), dB[l].data_ptr()
        lib = _bind()
        nbytes = ctypes.c_size_t()
        _native.check(lib.crdpn_pose_tail_backward_workspace_bytes(arr, len(spec), B, ctypes.byref(nbytes)),
                      "crdpn_pose_tail_backward_workspace_bytes")
        ws = torch.empty(nbytes.value, dtype=torch.uint8, device=dev)
        g_sf = torch.empty_like(sf) if ctx.needs_input_grad[1] else None
        g_img = torch.empty_like(img) if ctx.needs_input_grad[2] else None
        with _native.on_device(dev):
            rc = lib.crdpn_pose_tail_backward(arr, len(spec), sf.data_ptr(), img.data_ptr(), B, tail.shape_dim, tail.img_dim,
                                              _ptr(g_sf), _ptr(g_img), ws.data_ptr(), ws.numel(), _native.stream_ptr(dev))
        _native.check(rc, "crdpn_pose_tail_backward")
        grads = []
        for kind, l in tail._param_order():
            if kind == "W":
                grads.append(dW[l])
            elif kind == "b":
                grads.append(db[l])
            elif kind == "G":
                grads.append(dG[l])
            elif kind == "B":
                grads.append(dB[l])
            elif kind == "HW":     # head h's rows of the concatenated head layer
                lo = sum(tail.head_sizes[:l]); grads.append(dW[4][lo:lo + tail.head_sizes[l]])
            else:
                lo = sum(tail.head_sizes[:l]); grads.append(db[4][lo:lo + tail.head_sizes[l]])
        return (None, g_sf, g_img, *grads)


class PoseTail(nn.Module):
    """The tail of ``PoseEstimator`` (``auxiliary/model.py:238-272``) with the reference's parameter names, so that a
    ``PoseEstimator`` ``state_dict`` loads with ``strict=False`` / key filtering (``auxiliary/utils.py:56-73``):
    ``deformNet.{conv1..4, bn1..3}``, ``fc_{cls,reg}_{azi,ele,inp}``, ``projector.{0,1,3,4,6}``.

    ``forward(shape_feature [B, Fs], img_feature [B, Fi])`` returns what ``PoseEstimator.forward`` returns after its encoders:
    ``([cls_azi, cls_ele, cls_inp, reg_azi, reg_ele, reg_inp], x, projector(img_feature))``.  ``.eval()``: BatchNorm folded, one
    launch, no gradient path.  ``.train()``: batch-statistics BatchNorm inside the same kernel, running statistics updated,
    gradients to both inputs and every parameter.  B <= 256 rows per call in train mode (the reference uses 160)."""

    def __init__(self, img_feature_dim=1024, shape_feature_dim=256, azi_classes=24, ele_classes=12, inp_classes=24,
                 precision: str = "fp32"):
        super().__init__()
        if precision not in ("fp32", "bf16"):
            raise ValueError("precision must be 'fp32' or 'bf16'")
        self.precision, self.eps = precision, 1e-5
        self.shape_dim, self.img_dim = shape_feature_dim, img_feature_dim
        self.projector = nn.Sequential(nn.Linear(img_feature_dim, 800), nn.BatchNorm1d(800), nn.ReLU(inplace=True),
                                       nn.Linear(800, 400), nn.BatchNorm1d(400), nn.ReLU(inplace=True), nn.Linear(400, 200))
        self.deformNet = DeformNet(bottleneck_size=shape_feature_dim + img_feature_dim)
        self.fc_cls_azi = nn.Linear(200, azi_classes)
        self.fc_cls_ele = nn.Linear(200, ele_classes)
        self.fc_cls_inp = nn.Linear(200, inp_classes)
        self.fc_reg_azi = nn.Linear(200, azi_classes)
        self.fc_reg_ele = nn.Linear(200, ele_classes)
        self.fc_reg_inp = nn.Linear(200, inp_classes)
        self.head_sizes = [azi_classes, ele_classes, inp_classes, azi_classes, ele_classes, inp_classes]
        C = shape_feature_dim + img_feature_dim
        self.spec = _tail_spec(shape_feature_dim, img_feature_dim, [C, C // 2, C // 4, 200], sum(self.head_sizes), [800, 400, 200])
        self._chain = None
        self._frozen = None
        self._frozen_key = None

    def _apply(self, fn, *a, **k):
        self._chain = self._frozen = self._frozen_key = None
        return super()._apply(fn, *a, **k)

    def train(self, mode: bool = True):
        self._frozen = self._frozen_key = None
        return super().train(mode)

    def _ensure_chain(self, device):
        if self._chain is None or self._chain.device != device:
            self._chain = _Chain(self.spec, self.shape_dim, self.img_dim, device)
        return self._chain

    def _layer_params(self):
        d, p = self.deformNet, self.projector
        heads_W = torch.cat([getattr(self, h).weight.detach() for h in HEADS], 0)
        heads_b = torch.cat([getattr(self, h).bias.detach() for h in HEADS], 0)
        lin = {0: (d.conv1.weight, d.conv1.bias), 1: (d.conv2.weight, d.conv2.bias), 2: (d.conv3.weight, d.conv3.bias),
               3: (d.conv4.weight, d.conv4.bias), 4: (heads_W, heads_b), 5: (p[0].weight, p[0].bias), 6: (p[3].weight, p[3].bias),
               7: (p[6].weight, p[6].bias)}
        bns = {0: d.bn1, 1: d.bn2, 2: d.bn3, 5: p[1], 6: p[4]}
        return lin, bns

    def _param_order(self):
        """(kind, layer) for every tensor handed to the autograd function, in order."""
        order = []
        for l in (0, 1, 2, 3, 5, 6, 7):
            order += [("W", l), ("b", l)]
        for l in (0, 1, 2, 5, 6):
            order += [("G", l), ("B", l)]
        for h in range(6):
            order += [("HW", h), ("Hb", h)]
        return order

    def _param_tensors(self):
        lin, bns = self._layer_params()
        out = []
        for kind, l in self._param_order():
            if kind == "W":
                out.append(lin[l][0])
            elif kind == "b":
                out.append(lin[l][1])
            elif kind == "G":
                out.append(bns[l].weight)
            elif kind == "B":
                out.append(bns[l].bias)
            elif kind == "HW":
                out.append(getattr(self, HEADS[l]).weight)
            else:
                out.append(getattr(self, HEADS[l]).bias)
        return out

    def forward(self, shape_feature, img_feature):
        _check_inputs(shape_feature, img_feature, self.shape_dim, self.img_dim, "PoseTail")
        if not self.training:
            key = tuple(p._version for p in self.parameters()) + tuple(b._version for b in self.buffers())
            if self._frozen is None or self._frozen_key != key:
                sd = {k: v.detach() for k, v in self.state_dict().items()}
                self._frozen = FrozenPoseTail(sd, dtype=torch.bfloat16 if self.precision == "bf16" else torch.float32).to(img_feature.device)
                self._frozen_key = key
            return self._frozen(shape_feature, img_feature)
        if img_feature.shape[0] > 256:
            raise RuntimeError("PoseTail train mode: at most 256 rows per call (BatchNorm couples the whole batch)")
        if img_feature.shape[0] < 2:
            raise ValueError("Expected more than 1 value per channel when training")   # nn.BatchNorm1d's own rule
        sf = shape_feature.to(torch.float32).contiguous()
        img = img_feature.to(torch.float32).contiguous()
        heads, x, p = _PoseTailTrainFunction.apply(self, sf, img, *self._param_tensors())
        return list(torch.split(heads, self.head_sizes, dim=1)), x, p
