"""Frozen-teacher tail of ``PoseEstimator`` -- everything between the two encoders and the losses, for the KD loop.

Reference: ``auxiliary/model.py:183-203`` (``DeformNet``: four 1x1 Conv1d on a length-1 sequence = four Linear layers, three
BatchNorm1d + ReLU, tanh) and ``model.py:238-272`` (``cat`` of the shape and image features, the six ``fc_*`` heads, the
``projector`` MLP); call site ``KD/common/base_class.py:363`` where the teacher is in ``.eval()`` and frozen
(``base_class.py:317``), so BatchNorm uses running statistics and nothing needs a gradient.

What this does with that: the ~27 launches of the eager tail (cat, view, 4 conv, 5 BN, 5 ReLU,
tanh, 6 + 3 Linear) become 15 -- BatchNorm folded into the weights, the concat replaced by a split-K pair of GEMMs, the six
heads concatenated into one [120, 200] GEMM -- captured in ONE CUDA graph per batch size (one launch per step).  The GEMMs
are M = 138 rows against 16 MB of weights, i.e. weight-streaming bound (about 3 us of HBM time): plain library GEMMs
(cuBLAS through ``torch.addmm``), which is the right tool for them; there is no hand-written kernel here and no gradient
path (a student-side tail would need one).  fp32 throughout: outputs equal the reference's to fp32 rounding.
"""
from __future__ import annotations

import torch
from torch import nn

HEADS = ("fc_cls_azi", "fc_cls_ele", "fc_cls_inp", "fc_reg_azi", "fc_reg_ele", "fc_reg_inp")


def _fold(W, b, bn_w, bn_b, mean, var, eps):
    s = bn_w / torch.sqrt(var + eps)
    return W * s[:, None], (b - mean) * s + bn_b


class FrozenPoseTail(nn.Module):
    """forward(shape_feature [B, Fs], img_feature [B, Fi]) -> ([cls_azi, cls_ele, cls_inp, reg_azi, reg_ele, reg_inp], x [B, 200],
    projector(img_feature) [B, 200]) -- the three values ``PoseEstimator.forward`` returns (``model.py:272``).

    Build it from a trained ``PoseEstimator``'s ``state_dict`` (``from_state_dict``); call ``refold()`` after loading new
    weights.  With ``graph=True`` (default) the launch sequence is captured once per batch size and replayed; the returned
    tensors are then the graph's static outputs and are overwritten by the next call."""

    def __init__(self, folded: dict, shape_dim: int, head_sizes, graph: bool = True, dtype=torch.float32):
        super().__init__()
        if dtype not in (torch.float32, torch.bfloat16):
            raise ValueError("FrozenPoseTail dtype must be float32 (reference parity) or bfloat16 (1e-2 tolerance mode)")
        self.dtype = dtype
        for k, v in folded.items():
            self.register_buffer(k, v.to(dtype).contiguous())
        self.shape_dim = shape_dim
        self.head_sizes = [int(h) for h in head_sizes]
        self.graph = graph
        self._graphs = {}
        self._source = None

    # -- construction ---------------------------------------------------------------------------------------
    @staticmethod
    def fold_state_dict(sd: dict, prefix: str = "", eps: float = 1e-5):
        g = lambda k: sd[prefix + k].detach().to(torch.float32)
        out = {}
        for n in (1, 2, 3):
            W, b = _fold(g(f"deformNet.conv{n}.weight")[:, :, 0], g(f"deformNet.conv{n}.bias"), g(f"deformNet.bn{n}.weight"),
                         g(f"deformNet.bn{n}.bias"), g(f"deformNet.bn{n}.running_mean"), g(f"deformNet.bn{n}.running_var"), eps)
            out[f"W{n}t"], out[f"b{n}"] = W.t().contiguous(), b     # stored transposed: addmm(bias, x, Wt)
        out["W4t"], out["b4"] = g("deformNet.conv4.weight")[:, :, 0].t().contiguous(), g("deformNet.conv4.bias")
        out["Wht"] = torch.cat([g(h + ".weight") for h in HEADS], 0).t().contiguous()
        out["bh"] = torch.cat([g(h + ".bias") for h in HEADS], 0)
        for i, (lin, bn) in enumerate(((0, 1), (3, 4))):
            W, b = _fold(g(f"projector.{lin}.weight"), g(f"projector.{lin}.bias"), g(f"projector.{bn}.weight"),
                         g(f"projector.{bn}.bias"), g(f"projector.{bn}.running_mean"), g(f"projector.{bn}.running_var"), eps)
            out[f"P{i + 1}t"], out[f"p{i + 1}"] = W.t().contiguous(), b
        out["P3t"], out["p3"] = g("projector.6.weight").t().contiguous(), g("projector.6.bias")
        head_sizes = [sd[prefix + h + ".weight"].shape[0] for h in HEADS]
        img_dim = out["P1t"].shape[0]
        return out, out["W1t"].shape[0] - img_dim, head_sizes

    @classmethod
    def from_state_dict(cls, sd: dict, prefix: str = "", graph: bool = True, dtype=torch.float32) -> "FrozenPoseTail":
        """dtype=torch.bfloat16 stores the folded weights and runs the GEMMs in bf16 (fp32 accumulate, tensor cores): the
        north_star's 1e-2 tolerance mode; the default float32 reproduces the reference to fp32 rounding."""
        folded, shape_dim, head_sizes = cls.fold_state_dict(sd, prefix)
        m = cls(folded, shape_dim, head_sizes, graph, dtype)
        m._source = (sd, prefix)
        return m

    def refold(self, sd: dict | None = None, prefix: str | None = None):
        """Re-derive the folded weights (after the teacher's weights changed) and drop the captured graphs."""
        if sd is None:
            sd, prefix = self._source
        folded, _, _ = self.fold_state_dict(sd, prefix or "")
        for k, v in folded.items():
            getattr(self, k).copy_(v.to(self.dtype))
        self._graphs = {}
        return self

    def _apply(self, fn, *a, **k):
        self._graphs = {}
        return super()._apply(fn, *a, **k)

    # -- forward --------------------------------------------------------------------------------------------
    def _run(self, sf, img):
        Fs = self.shape_dim
        h = torch.addmm(self.b1, sf, self.W1t[:Fs])          # cat((shape, img), 1) @ W1^T as two split-K GEMMs
        h.addmm_(img, self.W1t[Fs:]).relu_()
        h = torch.addmm(self.b2, h, self.W2t).relu_()
        h = torch.addmm(self.b3, h, self.W3t).relu_()
        x = torch.addmm(self.b4, h, self.W4t).tanh_()
        heads = torch.addmm(self.bh, x, self.Wht)             # six heads in one GEMM
        p = torch.addmm(self.p1, img, self.P1t).relu_()
        p = torch.addmm(self.p2, p, self.P2t).relu_()
        p = torch.addmm(self.p3, p, self.P3t)
        return heads, x, p

    def _split(self, heads):
        return list(torch.split(heads, self.head_sizes, dim=1))

    @torch.no_grad()
    def forward(self, shape_feature, img_feature):
        if not (shape_feature.is_cuda and img_feature.is_cuda):
            raise RuntimeError("FrozenPoseTail inputs must be CUDA tensors: this package has no CPU fallback")
        sf, img = shape_feature.detach().to(self.dtype).contiguous(), img_feature.detach().to(self.dtype).contiguous()
        if not self.graph:
            heads, x, p = self._run(sf, img)
            return self._split(heads.float()), x.float(), p.float()
        key = (sf.shape[0], sf.device)
        entry = self._graphs.get(key)
        if entry is None:
            s_sf, s_img = torch.empty_like(sf), torch.empty_like(img)
            s_sf.copy_(sf); s_img.copy_(img)
            side = torch.cuda.Stream(device=sf.device)
            side.wait_stream(torch.cuda.current_stream(sf.device))
            with torch.cuda.stream(side):                     # warm-up outside capture (cuBLAS workspaces, autotuning)
                for _ in range(2):
                    self._run(s_sf, s_img)
            torch.cuda.current_stream(sf.device).wait_stream(side)
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph):
                heads, x, p = self._run(s_sf, s_img)
                outs = (heads.float(), x.float(), p.float())  # no-ops in float32 mode
            entry = self._graphs[key] = (graph, s_sf, s_img, outs)
        graph, s_sf, s_img, (heads, x, p) = entry
        s_sf.copy_(sf)
        s_img.copy_(img)
        graph.replay()
        return self._split(heads), x, p
