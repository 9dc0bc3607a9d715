"""oracle/kd_losses_oracle.py -- CPU restatement of the reference's in-batch contrastive KD losses and KD loss mixer.
TEST INFRASTRUCTURE ONLY (tests/, __graft_entry__.smoke(), bench.py's cpu_baseline leg); nothing in the package
imports it.

Restates, in plain torch ops (fp64 by default, differentiable so that autograd yields the oracle gradients):

* ``rotation_err``        auxiliary/utils.py:156-202 (``angles_to_matrix`` + geodesic angle, degrees)
* ``nce_kd``              auxiliary/model_utils.py:225-261 (``poseNCE_KD``) and :263-285 (``infoNCE_KD``: dropout(p=0.3,
                          training=True) on the teacher side, then the same NCE with all weights 1)
* ``nce_self`` / ``single_nce_kd`` / ``multipose_nce_kd``   auxiliary/model_utils.py:169-223, 288-351 (``infoNCE``, ``poseNCE``,
                          ``singleinfoNCE_KD``, ``multiposeNCE_KD``)
* ``kl_div_t``            KD/vision/vanilla/vanilla_kd.py:8-32 (``TemperatureScaledKLDivLoss``)
* ``ce_loss``/``delta_loss``  auxiliary/loss.py:7-34 (``CELoss``, ``DeltaLoss``: bin classification + SmoothL1 on tanh deltas)
* ``kd_loss_new``         KD/vision/vanilla/vanilla_kd.py:143-164 (``calculate_kd_loss_new``)
* ``student_kd_step_loss``  the loss part of one student step, KD/common/base_class.py:365-387

PINNED: tests/test_oracle_kd_losses.py checks every function against the reference's own code, imported unmodified
from /root/reference in the build container (matplotlib / pymesh stubbed, ``F.dropout`` replaced by the fixed keep-mask
below so both sides see the same mask), and against tests/golden/kd_losses_golden.npz produced by the reference
(oracle/gen_golden_kd.py).

The dropout keep-mask is this build's own counter-based stream (the reference uses torch's global RNG, which cannot be
reproduced across devices): element e = n*C + c takes word e%4 of Philox4x32-10 block (seed, offset + e/4);
u = (word >> 8) * 2^-24; keep iff u >= p; kept values are scaled by 1/(1-p).
"""
from __future__ import annotations

import math

import numpy as np
import torch

WEIGHTINGS = ("none", "linear", "square", "sqrt", "sin", "sinsin")


def philox_keep_mask(seed: int, offset: int, n_elem: int, p: float) -> np.ndarray:
    """bool[n_elem] keep-mask of the build's dropout stream (see module docstring)."""
    from oracle import crd_oracle
    out = np.empty(n_elem, dtype=bool)
    thr = np.float32(p)
    for blk in range((n_elem + 3) // 4):
        r = crd_oracle.philox(seed, offset + blk)
        for w in range(4):
            e = blk * 4 + w
            if e < n_elem:
                u = np.float32(int(r[w]) >> 8) * np.float32(2.0 ** -24)
                out[e] = u >= thr
    return out


def angles_to_matrix(angles: torch.Tensor) -> torch.Tensor:
    """R = Rz(inp) Rx(ele - pi/2) Rz(-azi) flattened to 9 columns (auxiliary/utils.py:156-178)."""
    azi, ele, rol = angles[:, 0], angles[:, 1], angles[:, 2]
    ca, sa, ce, se, cr, sr = torch.cos(azi), torch.sin(azi), torch.cos(ele), torch.sin(ele), torch.cos(rol), torch.sin(rol)
    return torch.stack((cr * ca - sr * ce * sa, sr * ca + cr * ce * sa, se * sa,
                        -cr * sa - sr * ce * ca, -sr * sa + cr * ce * ca, se * ca,
                        sr * se, -cr * se, ce), dim=1)


def rotation_err(preds: torch.Tensor, targets: torch.Tensor, dtype=torch.float64) -> torch.Tensor:
    """Geodesic angle in degrees between the rotations of two [n,3] (azi, ele, inp) label sets (utils.py:181-202)."""
    def prep(t):
        t = t.to(dtype).clone()
        t[:, 1] -= 180.0
        t[:, 2] -= 180.0
        return t * math.pi / 180.0
    Rp, Rg = angles_to_matrix(prep(preds)), angles_to_matrix(prep(targets))
    return torch.acos(((Rp * Rg).sum(1).clamp(-1.0, 3.0) - 1.0) / 2.0) * 180.0 / math.pi


def pose_weights(label: torch.Tensor, weighting: str, dtype=torch.float64) -> torch.Tensor:
    """[b,b] pairwise pose-distance weights of poseNCE_KD (model_utils.py:233-249)."""
    b = label.shape[0]
    lo = label.reshape(-1, 1, 3).repeat(1, b, 1).reshape(-1, 3)
    la = label.reshape(1, -1, 3).repeat(b, 1, 1).reshape(-1, 3)
    d = rotation_err(lo, la, dtype).reshape(b, b) / 180.0
    if weighting == "linear":
        return d
    if weighting == "square":
        return d ** 2
    if weighting == "sqrt":
        return torch.sqrt(d)
    if weighting == "sin":
        return torch.abs(torch.sin(d * math.pi))
    if weighting == "sinsin":
        return torch.sin(d * math.pi) ** 2
    raise ValueError(weighting)


def normalize(x: torch.Tensor) -> torch.Tensor:
    """F.normalize(x, dim=-1): x / max(||x||_2, 1e-12)."""
    return x / x.norm(dim=-1, keepdim=True).clamp_min(1e-12)


def nce_kd(feat_ori, feat_pos, label=None, tau=0.1, weighting="none", keep_mask=None, dropout_p=0.0, dtype=torch.float64):
    """poseNCE_KD (weighting != "none", label given) / infoNCE_KD (weighting == "none", keep_mask = the dropout mask).

    loss = mean_n -log( exp(a_n.p_n/tau) / (exp(a_n.p_n/tau) + sum_k exp(a_n.p_k/tau) * w_nk) ), the sum running over
    ALL k including k = n (model_utils.py:252-261)."""
    a = feat_ori.to(dtype)
    p = feat_pos.to(dtype)
    if keep_mask is not None:
        p = p * torch.as_tensor(keep_mask, dtype=dtype).reshape(p.shape) / (1.0 - dropout_p)
    a, p = normalize(a), normalize(p)
    l_pos = torch.exp((a * p).sum(1, keepdim=True) / tau)
    l_neg = torch.exp(a @ p.t() / tau)
    if weighting != "none":
        l_neg = l_neg * pose_weights(label, weighting, dtype)
    logits = torch.cat([l_pos, l_neg], dim=1)
    return (-torch.log(logits[:, 0] / logits.sum(-1))).mean()


def kl_div_t(y_pred, y, temperature=1.0, dtype=torch.float64):
    """T^2 * KLDivLoss(batchmean)(log_softmax(y_pred/T), softmax(y/T)) (vanilla_kd.py:22-32)."""
    log_p = torch.log_softmax(y_pred.to(dtype) / temperature, dim=1)
    q = torch.softmax(y.to(dtype) / temperature, dim=1)
    return temperature ** 2 * (torch.xlogy(q, q) - q * log_p).sum() / y_pred.shape[0]


def ce_loss(pred, target, rng, dtype=torch.float64):
    """CELoss(range)(pred, target): bin = range // n_classes; cross-entropy against target // bin (loss.py:7-20)."""
    bin_size = rng // pred.shape[1]
    return torch.nn.functional.cross_entropy(pred.to(dtype), (target // bin_size).long())


def delta_loss(pred_azi, pred_ele, pred_rol, target, bin_size, dtype=torch.float64):
    """DeltaLoss(bin)(...): SmoothL1(5 * tanh(pred[gt bin]) / 2, 5 * ((target % bin) / bin - 0.5)), mean over n*3 (loss.py:23-34)."""
    target = target.to(dtype)
    tdelta = (target % bin_size) / bin_size - 0.5
    tl = (target // bin_size).long()
    n = pred_azi.shape[0]
    ar = torch.arange(n)
    pd = torch.stack((pred_azi.to(dtype)[ar, tl[:, 0]].tanh() / 2, pred_ele.to(dtype)[ar, tl[:, 1]].tanh() / 2,
                      pred_rol.to(dtype)[ar, tl[:, 2]].tanh() / 2), dim=1)
    return torch.nn.functional.smooth_l1_loss(5.0 * pd, 5.0 * tdelta)


def kd_loss_new(y_pred_student, y_pred_teacher, student_features, teacher_features, gt_loss, temperature=1.0,
                w_gt=0.25, w_kl=0.75, w_rep=0.75, dtype=torch.float64):
    """calculate_kd_loss_new (vanilla_kd.py:143-164)."""
    kl = sum(kl_div_t(s, t, temperature, dtype) for s, t in zip(y_pred_student, y_pred_teacher))
    rep = kl_div_t(student_features, teacher_features, temperature, dtype)
    return w_kl * kl + w_gt * gt_loss + w_rep * rep


def gt_loss(out, label, bin_size=15, dtype=torch.float64):
    """loss_azi + loss_ele + loss_inp + loss_reg of one step (KD/common/base_class.py:365-369; training.py:50-54)."""
    return (ce_loss(out[0], label[:, 0], 360, dtype) + ce_loss(out[1], label[:, 1], 180, dtype) +
            ce_loss(out[2], label[:, 2], 360, dtype) + delta_loss(out[3], out[4], out[5], label.to(dtype), bin_size, dtype))


def student_kd_step_loss(out, teacher_out, student_features, teacher_features, label, bin_size=15, temperature=1.0,
                         dtype=torch.float64):
    """The loss of one student KD step (base_class.py:365-387): gt losses, then calculate_kd_loss_new."""
    return kd_loss_new(out, teacher_out, student_features, teacher_features, gt_loss(out, label, bin_size, dtype),
                       temperature, dtype=dtype)


# ---------------------------------------------------------------------------------------------------------
# ---- the other in-batch variants of auxiliary/model_utils.py (negatives from the anchors themselves; positive only;
# ---- several positives per anchor) -----------------------------------------------------------------------------------
def nce_self(feat_ori, feat_pos, label=None, tau=0.1, weighting="none", dtype=torch.float64, weights=None):
    """infoNCE (model_utils.py:169-186: cross-entropy of row n over [a_n.a_k/tau for k != n, a_n.p_n/tau at k = n]) and
    poseNCE (:189-223: the same with pose weights on the a_n.a_k terms; the k = n weight is f(0) = 0):
    loss = mean_n -log( e^{a_n.p_n/tau} / (e^{a_n.p_n/tau} + sum_{k != n} w_nk e^{a_n.a_k/tau}) ).
    `weights` overrides the [b,b] weight matrix (the tests substitute the reference's own fp32 distances)."""
    a, p = normalize(feat_ori.to(dtype)), normalize(feat_pos.to(dtype))
    b = a.shape[0]
    if weights is not None:
        w = weights.to(dtype)
    elif weighting == "none":
        w = 1.0 - torch.eye(b, dtype=dtype)
    else:
        w = pose_weights(label, weighting, dtype) * (1.0 - torch.eye(b, dtype=dtype))
    l_pos = torch.exp((a * p).sum(1) / tau)
    l_neg = (torch.exp(a @ a.t() / tau) * w).sum(1)
    return (-torch.log(l_pos / (l_pos + l_neg))).mean()


def single_nce_kd(feat_ori, feat_pos, tau=0.1, dtype=torch.float64):
    """singleinfoNCE_KD (model_utils.py:288-304): -log(exp(a_n.p_n/tau)) = -(a_n.p_n)/tau, averaged."""
    a, p = normalize(feat_ori.to(dtype)), normalize(feat_pos.to(dtype))
    return (-(a * p).sum(1) / tau).mean()


def pairwise_rotation_err(label, dtype=torch.float64):
    b = label.shape[0]
    lo = label.reshape(-1, 1, 3).repeat(1, b, 1).reshape(-1, 3)
    la = label.reshape(1, -1, 3).repeat(b, 1, 1).reshape(-1, 3)
    return rotation_err(lo, la, dtype).reshape(b, b)


def multipose_nce_kd(feat_ori, feat_pos, label, tau=0.1, threshold=30.0, dtype=torch.float64):
    """multiposeNCE_KD (model_utils.py:307-351): positives of anchor n = {k : k = n or rotation_err(n, k) <= 30 degrees};
    loss = mean_n -log( P_n / (P_n + sum_k e^{a_n.p_k/tau}) ), P_n = sum over the positives of e^{a_n.p_k/tau}."""
    a, p = normalize(feat_ori.to(dtype)), normalize(feat_pos.to(dtype))
    b = a.shape[0]
    mark = ((pairwise_rotation_err(label, dtype) <= threshold) | torch.eye(b, dtype=torch.bool)).to(dtype)
    E = torch.exp(a @ p.t() / tau)
    l_pos = (E * mark).sum(1)
    return (-torch.log(l_pos / (l_pos + E.sum(1)))).mean()


def clustered_labels(label, seed=3):
    """`label` with rows 1-3 moved to within a few degrees of row 0 and row 5 next to row 4, so that multiposeNCE_KD's
    30-degree rule finds several positives per anchor (random labels almost never fall that close)."""
    g = torch.Generator().manual_seed(seed)
    out = label.clone().float()
    for dst, src in ((1, 0), (2, 0), (3, 0), (5, 4)):
        out[dst] = out[src] + (torch.rand(3, generator=g) - 0.5) * 16.0
    out[:, 0] = out[:, 0] % 360.0
    out[:, 1] = out[:, 1].clamp(1.0, 179.0)
    out[:, 2] = out[:, 2] % 360.0
    return out


class cuda_is_identity:
    """Context manager: ``Tensor.cuda()`` returns the tensor itself, so that the reference functions that move their
    constants to the GPU (infoNCE's labels, multiposeNCE_KD's mark matrix) run unmodified on CPU tensors."""

    def __enter__(self):
        self.orig = torch.Tensor.cuda
        torch.Tensor.cuda = lambda self_, *a, **k: self_
        return self

    def __exit__(self, *a):
        torch.Tensor.cuda = self.orig
        return False


def load_reference(root: str = "/root/reference"):
    """Import the reference's own loss code (build container only).  Returns a namespace or None."""
    import sys
    import types
    from pathlib import Path
    if not Path(root).exists():
        return None
    for name in ("matplotlib", "matplotlib.pyplot", "pymesh"):
        if name not in sys.modules:
            sys.modules[name] = types.ModuleType(name)
    sys.modules["matplotlib"].use = lambda *a, **k: None
    sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]
    if root not in sys.path:
        sys.path.insert(0, root)
    from auxiliary import loss as ref_loss  # type: ignore
    from auxiliary import model_utils as ref_mu  # type: ignore
    from auxiliary import utils as ref_utils  # type: ignore
    from KD.vision.vanilla import vanilla_kd as ref_vkd  # type: ignore
    return types.SimpleNamespace(loss=ref_loss, model_utils=ref_mu, utils=ref_utils, vanilla_kd=ref_vkd)


class fixed_dropout:
    """Context manager: the reference's ``F.dropout`` applies ``keep_mask`` (scaled by 1/(1-p)) instead of torch's RNG."""

    def __init__(self, ref, keep_mask):
        self.F = ref.model_utils.F
        self.mask = keep_mask

    def __enter__(self):
        self.orig = self.F.dropout
        mask = self.mask

        def dropout(x, p=0.5, training=True, inplace=False):
            return x * torch.as_tensor(mask, dtype=x.dtype).reshape(x.shape) / (1.0 - p)
        self.F.dropout = dropout
        return self

    def __exit__(self, *a):
        self.F.dropout = self.orig
        return False


def synthetic_step(n=12, C=200, seed=46, bin_size=15):
    """Seeded synthetic tensors of one KD step: 6 student / teacher head outputs, features, integer labels."""
    g = torch.Generator().manual_seed(seed)
    widths = (360 // bin_size, 180 // bin_size, 360 // bin_size) * 2
    out = [torch.randn(n, w, generator=g) * 2 for w in widths]
    tout = [torch.randn(n, w, generator=g) * 2 for w in widths]
    sf = torch.randn(n, C, generator=g)
    tf = torch.randn(n, C, generator=g) + 0.5 * sf
    label = torch.stack((torch.randint(0, 360, (n,), generator=g), torch.randint(0, 180, (n,), generator=g),
                         torch.randint(0, 360, (n,), generator=g)), dim=1)
    return out, tout, sf, tf, label
