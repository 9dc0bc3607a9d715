"""oracle/gen_golden_kd.py -- make tests/golden/kd_losses_golden.npz from the REFERENCE's own loss code.

Run in the build container only (needs /root/reference):   python oracle/gen_golden_kd.py
The reference functions (auxiliary/model_utils.py:225-285 infoNCE_KD / poseNCE_KD, auxiliary/loss.py CELoss / DeltaLoss,
KD/vision/vanilla/vanilla_kd.py TemperatureScaledKLDivLoss / calculate_kd_loss_new) are imported unmodified, fed the
seeded synthetic step of oracle/kd_losses_oracle.synthetic_step, and their fp32 values and input gradients are stored.
infoNCE_KD's dropout is replaced by the build's fixed keep-mask (stored too).  TEST INFRASTRUCTURE ONLY.
"""
from __future__ import annotations

import sys
import types
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from oracle import kd_losses_oracle as ko  # noqa: E402

SEED, OFFSET, P_DROP = 46, 7, 0.3


def reference_step_loss(ref, out, tout, sf, tf, label, bin_size=15):
    """Body of the reference's student step loss, KD/common/base_class.py:365-387, with the reference's own classes."""
    crit_azi, crit_ele, crit_inp = ref.loss.CELoss(360), ref.loss.CELoss(180), ref.loss.CELoss(360)
    crit_reg = ref.loss.DeltaLoss(bin_size)
    gt = (crit_azi(out[0], label[:, 0]) + crit_ele(out[1], label[:, 1]) + crit_inp(out[2], label[:, 2]) +
          crit_reg(out[3], out[4], out[5], label.float()))
    self_ns = types.SimpleNamespace(loss_fn=ref.vanilla_kd.TemperatureScaledKLDivLoss(temperature=1.0))
    return ref.vanilla_kd.VanillaKD.calculate_kd_loss_new(self_ns, out, tout, sf, tf, gt), gt


def main(out_path: Path = ROOT / "tests" / "golden" / "kd_losses_golden.npz") -> None:
    ref = ko.load_reference()
    if ref is None:
        raise SystemExit("/root/reference is not mounted; golden vectors can only be made in the build container")
    torch.set_num_threads(4)
    n, C = 12, 200
    out, tout, sf, tf, label = ko.synthetic_step(n, C, seed=46)
    blob = {"n": np.int64(n), "C": np.int64(C), "seed": np.int64(SEED), "offset": np.int64(OFFSET), "p_drop": np.float64(P_DROP)}
    keep = ko.philox_keep_mask(SEED, OFFSET, n * C, P_DROP)
    blob["keep_mask"] = keep

    def leaf(t):
        return t.clone().requires_grad_()

    # infoNCE_KD (tau 0.1 as in training.py:57, and the KD default 0.5), both gradient directions
    for tau in (0.1, 0.5):
        a, p = leaf(sf), leaf(tf)
        with ko.fixed_dropout(ref, keep):
            l = ref.model_utils.infoNCE_KD(a, p, label, tau)
        l.backward()
        blob[f"infonce_kd/tau{tau}/loss"] = l.detach().numpy()
        blob[f"infonce_kd/tau{tau}/d_ori"] = a.grad.numpy()
        blob[f"infonce_kd/tau{tau}/d_pos"] = p.grad.numpy()
    for wt in ("linear", "square", "sqrt", "sin", "sinsin"):
        a, p = leaf(sf), leaf(tf)
        l = ref.model_utils.poseNCE_KD(a, p, label, 0.1, wt)
        l.backward()
        blob[f"posence_kd/{wt}/loss"] = l.detach().numpy()
        blob[f"posence_kd/{wt}/d_ori"] = a.grad.numpy()
        blob[f"posence_kd/{wt}/d_pos"] = p.grad.numpy()
    # the other in-batch variants (model_utils.py:169-223, 288-351); .cuda() inside them is made a no-op
    with ko.cuda_is_identity():
        a, p = leaf(sf), leaf(tf)
        l = ref.model_utils.infoNCE(a, p, 0.1)
        l.backward()
        blob["infonce/loss"], blob["infonce/d_ori"], blob["infonce/d_pos"] = l.detach().numpy(), a.grad.numpy(), p.grad.numpy()
        for wt in ("linear", "square", "sinsin"):
            a, p = leaf(sf), leaf(tf)
            l = ref.model_utils.poseNCE(a, p, label, 0.1, wt)
            l.backward()
            blob[f"posence/{wt}/loss"], blob[f"posence/{wt}/d_ori"], blob[f"posence/{wt}/d_pos"] = l.detach().numpy(), a.grad.numpy(), p.grad.numpy()
        a, p = leaf(sf), leaf(tf)
        l = ref.model_utils.singleinfoNCE_KD(a, p, label, 0.1)
        l.backward()
        blob["single/loss"], blob["single/d_ori"], blob["single/d_pos"] = l.detach().numpy(), a.grad.numpy(), p.grad.numpy()
        lab_c = ko.clustered_labels(label)
        a, p = leaf(sf), leaf(tf)
        l = ref.model_utils.multiposeNCE_KD(a, p, lab_c, 0.1)
        l.backward()
        blob["multipose/label"] = lab_c.numpy()
        blob["multipose/loss"], blob["multipose/d_ori"], blob["multipose/d_pos"] = l.detach().numpy(), a.grad.numpy(), p.grad.numpy()
    blob["rotation_err_pairs"] = ref.utils.rotation_err(label.reshape(-1, 1, 3).repeat(1, n, 1).reshape(-1, 3),
                                                        label.reshape(1, -1, 3).repeat(n, 1, 1).reshape(-1, 3)).numpy()
    # KL (temperatures 1 and 2), CE, Delta
    for T in (1.0, 2.0):
        s, t = leaf(out[0]), leaf(tout[0])
        l = ref.vanilla_kd.TemperatureScaledKLDivLoss(T)(s, t)
        l.backward()
        blob[f"kl/T{T}/loss"], blob[f"kl/T{T}/d_student"], blob[f"kl/T{T}/d_teacher"] = l.detach().numpy(), s.grad.numpy(), t.grad.numpy()
    s = leaf(out[1])
    l = ref.loss.CELoss(180)(s, label[:, 1])
    l.backward()
    blob["ce180/loss"], blob["ce180/d_pred"] = l.detach().numpy(), s.grad.numpy()
    d = [leaf(out[3]), leaf(out[4]), leaf(out[5])]
    l = ref.loss.DeltaLoss(15)(d[0], d[1], d[2], label.float())
    l.backward()
    blob["delta/loss"] = l.detach().numpy()
    for i in range(3):
        blob[f"delta/d_pred{i}"] = d[i].grad.numpy()
    # the whole student step loss
    o, to, a, p = [leaf(t) for t in out], [leaf(t) for t in tout], leaf(sf), leaf(tf)
    l, gt = reference_step_loss(ref, o, to, a, p, label)
    l.backward()
    blob["step/loss"], blob["step/gt_loss"] = l.detach().numpy(), gt.detach().numpy()
    for i in range(6):
        blob[f"step/d_out{i}"], blob[f"step/d_tout{i}"] = o[i].grad.numpy(), to[i].grad.numpy()
    blob["step/d_sf"], blob["step/d_tf"] = a.grad.numpy(), p.grad.numpy()
    for i in range(6):
        blob[f"in/out{i}"], blob[f"in/tout{i}"] = out[i].numpy(), tout[i].numpy()
    blob["in/sf"], blob["in/tf"], blob["in/label"] = sf.numpy(), tf.numpy(), label.numpy()
    out_path.parent.mkdir(parents=True, exist_ok=True)
    np.savez_compressed(out_path, **blob)
    print(f"wrote {out_path} ({out_path.stat().st_size/1e3:.1f} kB); step loss {float(l):.6f}")


if __name__ == "__main__":
    main()
