"""oracle/gen_golden.py -- make tests/golden/pointnet_golden.npz from the REFERENCE's own ShapeEncoderPC.

Run in the build container only (needs /root/reference):   python oracle/gen_golden.py
The reference module (auxiliary/model.py:154-180) is imported unmodified, fed seeded synthetic clouds and
a seeded state, and its eval-mode output, train-mode output, updated running statistics and parameter
gradients are stored (fp32).  The GPU box has no /root/reference; tests there read this file.
TEST INFRASTRUCTURE ONLY.
"""
from __future__ import annotations

import sys
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from oracle import pointnet_oracle as po  # noqa: E402


def main(out_path: Path = ROOT / "tests" / "golden" / "pointnet_golden.npz") -> None:
    torch.manual_seed(46)
    torch.set_num_threads(8)
    F, B, P = 1024, 3, 333  # ragged P: not a multiple of 8/16/128
    st = po.random_state(F, seed=46)
    x = po.random_clouds(B, P, seed=47)
    gout = torch.randn(B, F, generator=torch.Generator().manual_seed(48))

    ref_eval = po.make_reference_module(st, F, training=False)
    if ref_eval is None:
        raise SystemExit("/root/reference is not mounted; golden vectors can only be made in the build container")
    with torch.no_grad():
        eval_out = ref_eval(x)

    ref_train = po.make_reference_module(st, F, training=True)
    train_out = ref_train(x)
    (train_out * gout).sum().backward()
    after = ref_train.state_dict()

    blob = {"x": x.numpy(), "gout": gout.numpy(), "eval_out": eval_out.numpy(),
            "train_out": train_out.detach().numpy(), "feature_dim": np.int64(F)}
    for k, v in st.items():
        blob["state/" + k] = v.numpy()
    for k, v in after.items():
        if "running" in k or "num_batches" in k:
            blob["after/" + k] = v.numpy()
    for k, p in ref_train.named_parameters():
        blob["grad/" + k] = p.grad.numpy()
    out_path.parent.mkdir(parents=True, exist_ok=True)
    np.savez_compressed(out_path, **blob)
    print(f"wrote {out_path} ({out_path.stat().st_size/1e6:.2f} MB); eval_out[0,:4]={eval_out[0,:4].tolist()}")


if __name__ == "__main__":
    main()
