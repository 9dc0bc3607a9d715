"""In-batch contrastive KD losses and the KD loss mixer -- host-side mirror of the reference's loss code on B200 kernels.

Same names, argument order and semantics as the reference:

    infoNCE_KD(feat_ori, feat_pos, label, tau=0.1, weighting="linear")      auxiliary/model_utils.py:263-285
    poseNCE_KD(feat_ori, feat_pos, label, tau=0.1, weighting="linear")      auxiliary/model_utils.py:225-261
    infoNCE(feat_ori, feat_pos, tau=0.1), poseNCE(feat_ori, feat_pos, label, tau, weighting)
                                                                            auxiliary/model_utils.py:169-223
    singleinfoNCE_KD(feat_ori, feat_pos, label, tau, weighting), multiposeNCE_KD(feat_ori, feat_pos, label, tau)
                                                                            auxiliary/model_utils.py:288-351
    CELoss(range)(pred, target), DeltaLoss(bin)(azi, ele, rol, target)      auxiliary/loss.py:7-34
    TemperatureScaledKLDivLoss(temperature)(y_pred, y)                      KD/vision/vanilla/vanilla_kd.py:8-32
    calculate_kd_loss_new(y_pred_student, y_pred_teacher, student_features, teacher_features, gt_loss)
                                                                            KD/vision/vanilla/vanilla_kd.py:143-164
plus ``student_kd_step_loss`` -- everything the student step adds up at KD/common/base_class.py:365-387 (three CE
terms, the delta regression term, seven temperature-scaled KL terms and their weights) in ONE launch, gradients in one
more.  The eager formulation needs ~50 forward and ~80 backward launches on [138, 24] tensors.

Every function is a ``torch.autograd.Function`` over the C ABI of ``libcrdpn_b200.so`` (``crdpn_nce_kd_*``,
``crdpn_kd_mix_*``); gradients are closed-form and deterministic.  No CPU path: tensors must be CUDA float32.

``infoNCE_KD`` drops 30 % of the teacher features unconditionally, as the reference does (``F.dropout(p=0.3,
training=True)``, model_utils.py:268).  The keep-mask comes from this package's counter-based Philox stream (seed =
``torch.initial_seed()`` unless ``set_dropout_stream`` is called; the offset advances by ceil(B*C/4) per call), so runs
are reproducible and the mask can be regenerated in backward instead of stored.
"""
from __future__ import annotations

import ctypes

import torch
from torch import nn

from . import _native

WEIGHTINGS = {"none": 0, "linear": 1, "square": 2, "sqrt": 3, "sin": 4, "sinsin": 5}
MODE_SELF, MODE_SINGLE, MODE_MULTI = 0x100, 0x200, 0x400   # crdpn_nce_kd_forward mode bits (include/crdpn_b200.h)
INFONCE_DROPOUT_P = 0.3

_stream_state = {"seed": None, "offset": 0}


def set_dropout_stream(seed: int, offset: int = 0) -> None:
    """Key and position of the Philox stream infoNCE_KD's dropout draws from."""
    _stream_state["seed"] = int(seed) & 0xFFFFFFFFFFFFFFFF
    _stream_state["offset"] = int(offset)


def _next_dropout_blocks(n_elem: int):
    if _stream_state["seed"] is None:
        _stream_state["seed"] = int(torch.initial_seed()) & 0xFFFFFFFFFFFFFFFF
    off = _stream_state["offset"]
    _stream_state["offset"] = off + (n_elem + 3) // 4
    return _stream_state["seed"], off


_stream_ptr = _native.stream_ptr


def _f32_cuda(t: torch.Tensor, name: str) -> torch.Tensor:
    if not t.is_cuda:
        raise RuntimeError(f"{name} must be a CUDA tensor: this package has no CPU fallback")
    if t.dtype != torch.float32:
        raise RuntimeError(f"{name} must be float32")
    return t.contiguous()


# ----------------------------------------------------------------------------------------------------------
class _NceKdFunction(torch.autograd.Function):
    """crdpn_nce_kd_forward (2 launches) / crdpn_nce_kd_backward (1 launch)."""

    @staticmethod
    def forward(ctx, feat_ori, feat_pos, label, tau, weighting, p_drop, seed, offset):
        a = _f32_cuda(feat_ori.detach(), "feat_ori")
        p = _f32_cuda(feat_pos.detach(), "feat_pos")
        if a.dim() != 2 or a.shape != p.shape:
            raise RuntimeError("feat_ori and feat_pos must both be [B, C]")
        B, C = a.shape
        dev = a.device
        lab = None
        if (weighting & 0xff) != 0 or (weighting & MODE_MULTI):
            lab = label.detach().to(device=dev, dtype=torch.float32).contiguous()
            if lab.shape != (B, 3):
                raise RuntimeError("label must be [B, 3] (azimuth, elevation, in-plane rotation in degrees)")
        lib = _native.lib()
        nbytes = ctypes.c_size_t(0)
        _native.check(lib.crdpn_nce_kd_workspace_bytes(B, C, ctypes.byref(nbytes)), "crdpn_nce_kd_workspace_bytes")
        ws = torch.empty(nbytes.value, dtype=torch.uint8, device=dev)
        loss = torch.empty((), dtype=torch.float32, device=dev)
        with _native.on_device(dev):
            rc = lib.crdpn_nce_kd_forward(a.data_ptr(), p.data_ptr(), lab.data_ptr() if lab is not None else None, B, C,
                                          float(tau), weighting, float(p_drop), seed, offset, loss.data_ptr(), ws.data_ptr(),
                                          ws.numel(), _stream_ptr(dev))
        _native.check(rc, "crdpn_nce_kd_forward")
        ctx.save_for_backward(ws)
        ctx.cfg = (B, C, float(tau), float(p_drop), seed, offset, feat_ori.shape, feat_pos.shape)
        ctx.need_pos = ctx.needs_input_grad[1]
        return loss

    @staticmethod
    def backward(ctx, grad_out):
        (ws,) = ctx.saved_tensors
        B, C, tau, p_drop, seed, offset, shp_a, shp_p = ctx.cfg
        dev = ws.device
        g = grad_out.detach().to(torch.float32).contiguous()
        d_ori = torch.empty(B, C, dtype=torch.float32, device=dev)
        d_pos = torch.empty(B, C, dtype=torch.float32, device=dev) if ctx.need_pos else None
        with _native.on_device(dev):
            rc = _native.lib().crdpn_nce_kd_backward(g.data_ptr(), B, C, tau, p_drop, seed, offset, ws.data_ptr(), ws.numel(),
                                                     d_ori.data_ptr(), d_pos.data_ptr() if d_pos is not None else None,
                                                     _stream_ptr(dev))
        _native.check(rc, "crdpn_nce_kd_backward")
        return d_ori.view(shp_a), (d_pos.view(shp_p) if d_pos is not None else None), None, None, None, None, None, None


def poseNCE_KD(feat_ori, feat_pos, label, tau=0.1, weighting="linear"):
    """Pose-weighted in-batch NCE between two embeddings (reference: auxiliary/model_utils.py:225-261)."""
    if weighting not in WEIGHTINGS or weighting == "none":
        # the reference silently keeps the raw degree distance for an unknown string; refuse instead of guessing
        raise ValueError(f"weighting must be one of {[w for w in WEIGHTINGS if w != 'none']}")
    return _NceKdFunction.apply(feat_ori, feat_pos, label, tau, WEIGHTINGS[weighting], 0.0, 0, 0)


def infoNCE_KD(feat_ori, feat_pos, label=None, tau=0.1, weighting="linear"):
    """In-batch InfoNCE between student and (dropped-out) teacher embeddings (reference: model_utils.py:263-285;
    ``label`` and ``weighting`` are accepted and ignored exactly as there)."""
    seed, offset = _next_dropout_blocks(feat_pos.numel())
    return _NceKdFunction.apply(feat_ori, feat_pos, None, tau, 0, INFONCE_DROPOUT_P, seed, offset)


def infoNCE(feat_ori, feat_pos, tau=0.1):
    """In-batch InfoNCE whose negatives are the OTHER anchors (reference: auxiliary/model_utils.py:169-186): cross-entropy of
    row n over logits [a_n.a_k / tau for k != n, a_n.q_n / tau at k = n]."""
    return _NceKdFunction.apply(feat_ori, feat_pos, None, tau, MODE_SELF, 0.0, 0, 0)


def poseNCE(feat_ori, feat_pos, label, tau=0.1, weighting="linear"):
    """Pose-weighted NCE whose negatives are the anchors themselves (reference: auxiliary/model_utils.py:189-223)."""
    if weighting not in WEIGHTINGS or weighting == "none":
        raise ValueError(f"weighting must be one of {[w for w in WEIGHTINGS if w != 'none']}")
    return _NceKdFunction.apply(feat_ori, feat_pos, label, tau, WEIGHTINGS[weighting] | MODE_SELF, 0.0, 0, 0)


def singleinfoNCE_KD(feat_ori, feat_pos, label=None, tau=0.1, weighting="linear"):
    """mean_n -(a_n.q_n) / tau on the normalised embeddings (reference: model_utils.py:288-304; ``label`` / ``weighting`` are
    accepted and ignored as there)."""
    return _NceKdFunction.apply(feat_ori, feat_pos, None, tau, MODE_SINGLE, 0.0, 0, 0)


def multiposeNCE_KD(feat_ori, feat_pos, label, tau=0.1):
    """NCE with several positives per anchor: every teacher row whose pose lies within 30 degrees of the anchor's (and the
    anchor's own) counts as a positive (reference: model_utils.py:307-351)."""
    return _NceKdFunction.apply(feat_ori, feat_pos, label, tau, MODE_MULTI, 0.0, 0, 0)


# ----------------------------------------------------------------------------------------------------------
_T_KL_HEAD, _T_KL_FEAT, _T_CE, _T_DELTA = 0, 6, 7, 10
_mix_ws = {}


def _mix_workspace(n: int, dev):
    """Zero-initialised once (the kernels' ticket word lives in it and every call leaves it zeroed)."""
    key = (dev.index if dev.index is not None else torch.cuda.current_device())
    ws = _mix_ws.get(key)
    if ws is None or ws.numel() < 4 * (n + 4):
        ws = _mix_ws[key] = torch.zeros(4 * (max(n, 1024) + 4), dtype=torch.uint8, device=dev)
    return ws


def _ptr_array(tensors):
    arr = (ctypes.c_void_p * 6)()
    for i, t in enumerate(tensors):
        arr[i] = t.data_ptr() if t is not None else None
    return arr


class _KdMixFunction(torch.autograd.Function):
    """crdpn_kd_mix_forward / crdpn_kd_mix_backward (one launch each).  Tensor arguments: 6 student head outputs,
    6 teacher head outputs, student features, teacher features (any may be None when its terms are off)."""

    @staticmethod
    def forward(ctx, cfg, label, *tensors):
        terms, T, w_kl, w_rep, w_gt, ce_bin, delta_bin = cfg
        ts = [(_f32_cuda(t.detach(), "loss input") if t is not None else None) for t in tensors]
        s_out, t_out, sf, tf = ts[0:6], ts[6:12], ts[12], ts[13]
        first = next(t for t in ts if t is not None)
        n, dev = first.shape[0], first.device
        widths = (ctypes.c_int32 * 6)(*[(t.shape[1] if t is not None else 0) for t in s_out])
        for a, b in list(zip(s_out, t_out)) + [(sf, tf)]:
            if a is not None and (a.dim() != 2 or a.shape[0] != n or (b is not None and b.shape != a.shape)):
                raise RuntimeError("student / teacher tensors must be [n, width] with matching shapes")
        lab, stride = None, 0
        if terms >> _T_CE:
            lab = label.detach().to(device=dev, dtype=torch.float32).contiguous()
            if lab.dim() == 1:
                lab = lab.view(n, 1)
            if lab.shape[0] != n:
                raise RuntimeError("label must have one row per sample")
            stride = lab.shape[1]
        ceb = (ctypes.c_int32 * 3)(*ce_bin)
        ws = _mix_workspace(n, dev)
        loss = torch.empty((), dtype=torch.float32, device=dev)
        args = (_ptr_array(s_out), _ptr_array(t_out), widths, sf.data_ptr() if sf is not None else None,
                tf.data_ptr() if tf is not None else None, sf.shape[1] if sf is not None else 0,
                lab.data_ptr() if lab is not None else None, stride, n, ceb, delta_bin, terms, T, w_kl, w_rep, w_gt)
        with _native.on_device(dev):
            rc = _native.lib().crdpn_kd_mix_forward(*args, loss.data_ptr(), ws.data_ptr(), ws.numel(), _stream_ptr(dev))
        _native.check(rc, "crdpn_kd_mix_forward")
        # the raw pointers in `args` stay valid as long as the saved tensors live; save_for_backward also makes autograd
        # refuse a backward after one of the inputs was modified in place (the backward kernel re-reads them)
        present = [i for i, t in enumerate(ts) if t is not None]
        ctx.save_for_backward(*[ts[i] for i in present], *([lab] if lab is not None else []))
        ctx.present = present
        ctx.args = args
        ctx.shapes = [(t.shape if t is not None else None) for t in tensors]
        return loss

    @staticmethod
    def backward(ctx, grad_out):
        saved = ctx.saved_tensors
        ts = [None] * 14
        for i, t in zip(ctx.present, saved):
            ts[i] = t
        dev = grad_out.device
        g = grad_out.detach().to(torch.float32).contiguous()
        need = ctx.needs_input_grad[2:]
        grads = [(torch.empty_like(t) if (t is not None and need[i]) else None) for i, t in enumerate(ts)]
        n = ctx.args[8]
        ws = _mix_workspace(n, dev)
        with _native.on_device(dev):
            rc = _native.lib().crdpn_kd_mix_backward(*ctx.args, g.data_ptr(), _ptr_array(grads[0:6]), _ptr_array(grads[6:12]),
                                                     grads[12].data_ptr() if grads[12] is not None else None,
                                                     grads[13].data_ptr() if grads[13] is not None else None,
                                                     ws.data_ptr(), ws.numel(), _stream_ptr(dev))
        _native.check(rc, "crdpn_kd_mix_backward")
        out = [(gr.view(shp) if gr is not None else None) for gr, shp in zip(grads, ctx.shapes)]
        return (None, None, *out)


def _mix(terms, label, s_out, t_out, sf, tf, T=1.0, w_kl=1.0, w_rep=1.0, w_gt=1.0, ce_bin=(1, 1, 1), delta_bin=1):
    cfg = (int(terms), float(T), float(w_kl), float(w_rep), float(w_gt), tuple(int(b) for b in ce_bin), int(delta_bin))
    return _KdMixFunction.apply(cfg, label, *s_out, *t_out, sf, tf)


_NONE6 = (None,) * 6


class TemperatureScaledKLDivLoss(nn.Module):
    """T^2 * KLDivLoss(batchmean)(log_softmax(y_pred / T), softmax(y / T)) (reference: vanilla_kd.py:8-32)."""

    def __init__(self, temperature):
        super().__init__()
        self.temperature = temperature

    def forward(self, y_pred, y):
        return _mix(1 << _T_KL_FEAT, None, _NONE6, _NONE6, y_pred, y, T=self.temperature)


class CELoss(nn.Module):
    """Cross-entropy on angle bins: bin size = range // n_classes, class = target // bin size (reference: loss.py:7-20)."""

    def __init__(self, range):
        super().__init__()
        self.__range__ = range

    def forward(self, pred, target):
        bin_size = self.__range__ // pred.size(1)
        return _mix(1 << _T_CE, target, (pred,) + (None,) * 5, _NONE6, None, None, ce_bin=(bin_size, 1, 1))


class DeltaLoss(nn.Module):
    """SmoothL1 between 5 * tanh(pred[gt bin]) / 2 and 5 * ((target % bin) / bin - 0.5) (reference: loss.py:23-44)."""

    def __init__(self, bin):
        super().__init__()
        self.__bin__ = bin

    def forward(self, pred_azi, pred_ele, pred_rol, target):
        return _mix(1 << _T_DELTA, target, (None, None, None, pred_azi, pred_ele, pred_rol), _NONE6, None, None,
                    delta_bin=self.__bin__)


def calculate_kd_loss_new(y_pred_student, y_pred_teacher, student_features, teacher_features, gt_loss, temperature=1.0):
    """0.25 * gt_loss + 0.75 * sum_i KL(student_i, teacher_i) + 0.75 * KL(student_features, teacher_features)
    (reference: VanillaKD.calculate_kd_loss_new, vanilla_kd.py:143-164, with its loss_fn = TemperatureScaledKLDivLoss(1.0),
    vanilla_kd.py:107).  The seven KL terms are one launch."""
    kl = _mix(0x7F, None, tuple(y_pred_student[:6]), tuple(y_pred_teacher[:6]), student_features, teacher_features,
              T=temperature, w_kl=0.75, w_rep=0.75)
    return kl + 0.25 * gt_loss


def student_kd_step_loss(out, teacher_out, student_features, teacher_features, label, bin_size=15, temperature=1.0):
    """The loss of one student KD step, KD/common/base_class.py:365-387, in one launch:
    0.25 * (CE_azi + CE_ele + CE_inp + Delta) + 0.75 * sum_{i<6} KL(out_i, teacher_out_i) + 0.75 * KL(features)."""
    ce_bin = (360 // out[0].size(1), 180 // out[1].size(1), 360 // out[2].size(1))
    return _mix(0x7FF, label, tuple(out[:6]), tuple(teacher_out[:6]), student_features, teacher_features, T=temperature,
                w_kl=0.75, w_rep=0.75, w_gt=0.25, ce_bin=ce_bin, delta_bin=bin_size)


def gt_loss(out, label, bin_size=15):
    """loss_azi + loss_ele + loss_inp + loss_reg (training.py:50-54; base_class.py:365-369) in one launch."""
    ce_bin = (360 // out[0].size(1), 180 // out[1].size(1), 360 // out[2].size(1))
    return _mix(0x780, label, tuple(out[:6]), _NONE6, None, None, ce_bin=ce_bin, delta_bin=bin_size)
