"""oracle/pose_tail_oracle.py -- CPU restatement of the reference's PoseEstimator tail (auxiliary/model.py:183-203, 238-272)
in eval mode: cat -> DeformNet (conv k=1 as matmul + explicit BatchNorm + ReLU, tanh) -> six heads, and projector(img).
TEST INFRASTRUCTURE ONLY.  PINNED: tests/test_oracle_pose_tail.py checks it against the reference's own modules (a small
PoseEstimator built from /root/reference in the build container) and against tests/golden/pose_tail_golden.npz made by
them (oracle/gen_golden_pose_tail.py)."""
from __future__ import annotations

import torch

HEADS = ("fc_cls_azi", "fc_cls_ele", "fc_cls_inp", "fc_reg_azi", "fc_reg_ele", "fc_reg_inp")


def _bn(x, sd, name, eps=1e-5):
    d = x.dtype
    return (x - sd[name + ".running_mean"].to(d)) / torch.sqrt(sd[name + ".running_var"].to(d) + eps) * sd[name + ".weight"].to(d) \
        + sd[name + ".bias"].to(d)


def forward(sd: dict, shape_feature, img_feature, dtype=torch.float64):
    """-> ([6 head outputs], x [B,200], projector(img_feature) [B,200]) as PoseEstimator.forward returns them (model.py:272)."""
    g = torch.cat((shape_feature.to(dtype), img_feature.to(dtype)), 1)             # model.py:260
    h = g
    for n in (1, 2, 3):                                                              # model.py:197-199
        h = torch.relu(_bn(h @ sd[f"deformNet.conv{n}.weight"].to(dtype)[:, :, 0].t() + sd[f"deformNet.conv{n}.bias"].to(dtype),
                           sd, f"deformNet.bn{n}"))
    x = torch.tanh(h @ sd["deformNet.conv4.weight"].to(dtype)[:, :, 0].t() + sd["deformNet.conv4.bias"].to(dtype))  # :200
    outs = [x @ sd[k + ".weight"].to(dtype).t() + sd[k + ".bias"].to(dtype) for k in HEADS]                        # :265-271
    p = img_feature.to(dtype)
    p = torch.relu(_bn(p @ sd["projector.0.weight"].to(dtype).t() + sd["projector.0.bias"].to(dtype), sd, "projector.1"))
    p = torch.relu(_bn(p @ sd["projector.3.weight"].to(dtype).t() + sd["projector.3.bias"].to(dtype), sd, "projector.4"))
    p = p @ sd["projector.6.weight"].to(dtype).t() + sd["projector.6.bias"].to(dtype)
    return outs, x, p


def tail_keys(sd: dict) -> dict:
    """The tail's entries of a PoseEstimator state_dict (encoders dropped)."""
    return {k: v for k, v in sd.items() if k.startswith(("deformNet.", "fc_", "projector."))}


def reference_tail(img_dim=64, shape_dim=32, seed=46, root="/root/reference"):
    """A small PoseEstimator from the reference (eval mode, BN statistics and affine randomised).  Returns (module, tail state
    dict) or None where the reference is not mounted."""
    import sys
    import types
    from pathlib import Path
    if not Path(root).exists():
        return None
    for name in ("matplotlib", "matplotlib.pyplot"):
        if name not in sys.modules:
            sys.modules[name] = types.ModuleType(name)
    sys.modules["matplotlib"].use = lambda *a, **k: None
    sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]
    if root not in sys.path:
        sys.path.insert(0, root)
    from auxiliary.model import PoseEstimator  # type: ignore
    torch.manual_seed(seed)
    m = PoseEstimator(img_feature_dim=img_dim, shape_feature_dim=shape_dim, shape="PointCloud").eval()
    g = torch.Generator().manual_seed(seed + 1)
    with torch.no_grad():
        for mod in list(m.deformNet.modules()) + list(m.projector.modules()):
            if isinstance(mod, torch.nn.BatchNorm1d):
                mod.weight.copy_(torch.randn(mod.weight.shape, generator=g))
                mod.bias.copy_(torch.randn(mod.bias.shape, generator=g))
                mod.running_mean.copy_(torch.randn(mod.running_mean.shape, generator=g) * 0.3)
                mod.running_var.copy_(torch.rand(mod.running_var.shape, generator=g) * 1.5 + 0.5)
    return m, tail_keys(m.state_dict())


def reference_forward(m, shape_feature, img_feature):
    """The reference's own modules on given encoder outputs: the body of PoseEstimator.forward after the encoders."""
    with torch.no_grad():
        gf = torch.cat((shape_feature, img_feature), 1)
        x = m.deformNet(gf.view(-1, gf.size(1), 1))
        outs = [getattr(m, k)(x) for k in HEADS]
        return outs, x, m.projector(img_feature)


# ---- train mode (training.py:30,47,75): batch-statistics BatchNorm, gradients by autograd -----------------------------
def _bn_train(x, sd, name, eps=1e-5):
    mean = x.mean(0)
    var = x.var(0, unbiased=False)
    return (x - mean) / torch.sqrt(var + eps) * sd[name + ".weight"] + sd[name + ".bias"], mean, x.var(0, unbiased=True)


def forward_train(sd: dict, shape_feature, img_feature, dtype=torch.float64, momentum=0.1):
    """Train-mode tail in `dtype` on tensors that may require grad (``sd`` values are used as given, so pass float64 leaves
    to differentiate).  -> (outs, x, p, new_running) where new_running maps '<bn>.running_mean/var' to the values after this
    step (momentum update with the unbiased batch variance, as nn.BatchNorm1d does)."""
    new_running = {}

    def bn(x, name):
        y, mean, uvar = _bn_train(x, sd, name)
        new_running[name + ".running_mean"] = (1 - momentum) * sd[name + ".running_mean"].to(dtype) + momentum * mean.detach()
        new_running[name + ".running_var"] = (1 - momentum) * sd[name + ".running_var"].to(dtype) + momentum * uvar.detach()
        return y

    h = torch.cat((shape_feature.to(dtype), img_feature.to(dtype)), 1)
    for n in (1, 2, 3):
        W = sd[f"deformNet.conv{n}.weight"]
        h = torch.relu(bn(h @ W[:, :, 0].t() + sd[f"deformNet.conv{n}.bias"], f"deformNet.bn{n}"))
    x = torch.tanh(h @ sd["deformNet.conv4.weight"][:, :, 0].t() + sd["deformNet.conv4.bias"])
    outs = [x @ sd[k + ".weight"].t() + sd[k + ".bias"] for k in HEADS]
    p = img_feature.to(dtype)
    p = torch.relu(bn(p @ sd["projector.0.weight"].t() + sd["projector.0.bias"], "projector.1"))
    p = torch.relu(bn(p @ sd["projector.3.weight"].t() + sd["projector.3.bias"], "projector.4"))
    p = p @ sd["projector.6.weight"].t() + sd["projector.6.bias"]
    return outs, x, p, new_running


def train_step_with_grads(sd: dict, shape_feature, img_feature, g_outs, g_x, g_p, dtype=torch.float64):
    """One train-mode forward + backward of L = sum <out_i, g_i> + <x, g_x> + <p, g_p>.  Returns (outs, x, p, new_running,
    grads) with grads keyed like the state dict plus 'in/shape_feature', 'in/img_feature'."""
    leaves = {k: v.detach().to(dtype).clone().requires_grad_(v.is_floating_point() and "running" not in k)
              for k, v in sd.items() if v.is_floating_point()}
    sf = shape_feature.detach().to(dtype).clone().requires_grad_(True)
    img = img_feature.detach().to(dtype).clone().requires_grad_(True)
    outs, x, p, new_running = forward_train(leaves, sf, img, dtype)
    loss = sum((o * g.to(dtype)).sum() for o, g in zip(outs, g_outs)) + (x * g_x.to(dtype)).sum() + (p * g_p.to(dtype)).sum()
    loss.backward()
    grads = {k: v.grad for k, v in leaves.items() if v.requires_grad and v.grad is not None}
    grads["in/shape_feature"], grads["in/img_feature"] = sf.grad, img.grad
    return [o.detach() for o in outs], x.detach(), p.detach(), new_running, grads


def reference_train_step(m, shape_feature, img_feature, g_outs, g_x, g_p):
    """The reference's own modules in train mode (fp32): the same forward + backward.  `m` is modified (running statistics,
    .grad); returns (outs, x, p, grads) keyed like the tail's state dict plus the two inputs."""
    m.train()
    for q in m.parameters():
        q.grad = None
    sf = shape_feature.clone().requires_grad_(True)
    img = img_feature.clone().requires_grad_(True)
    gf = torch.cat((sf, img), 1)
    x = m.deformNet(gf.view(-1, gf.size(1), 1))
    outs = [getattr(m, k)(x) for k in HEADS]
    p = m.projector(img)
    loss = sum((o * g).sum() for o, g in zip(outs, g_outs)) + (x * g_x).sum() + (p * g_p).sum()
    loss.backward()
    grads = {k: v.grad.detach().clone() for k, v in m.named_parameters()
             if k.startswith(("deformNet.", "fc_", "projector.")) and v.grad is not None}
    grads["in/shape_feature"], grads["in/img_feature"] = sf.grad.detach().clone(), img.grad.detach().clone()
    return [o.detach() for o in outs], x.detach(), p.detach(), grads
