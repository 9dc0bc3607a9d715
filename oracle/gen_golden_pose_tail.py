"""oracle/gen_golden_pose_tail.py -- make tests/golden/pose_tail_golden.npz from the REFERENCE's own PoseEstimator modules
(auxiliary/model.py:183-203, 238-272; a small instance: image feature 64, shape feature 32).  Build container only."""
from __future__ import annotations

import sys
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from oracle import pose_tail_oracle as pto  # noqa: E402


def main(out_path: Path = ROOT / "tests" / "golden" / "pose_tail_golden.npz") -> None:
    r = pto.reference_tail()
    if r is None:
        raise SystemExit("/root/reference is not mounted; golden vectors can only be made in the build container")
    m, sd = r
    g = torch.Generator().manual_seed(7)
    sf, img = torch.randn(9, 32, generator=g), torch.randn(9, 64, generator=g)
    outs, x, p = pto.reference_forward(m, sf, img)
    blob = {"in/shape_feature": sf.numpy(), "in/img_feature": img.numpy(), "out/x": x.numpy(), "out/projector": p.numpy()}
    for i, o in enumerate(outs):
        blob[f"out/head{i}"] = o.numpy()
    for k, v in sd.items():
        blob["state/" + k] = v.numpy()
    # train mode (training.py:30,47,75): batch-statistics BatchNorm, one forward + backward with fixed upstream gradients
    import copy
    mt = copy.deepcopy(m)
    gg = torch.Generator().manual_seed(8)
    g_outs = [torch.randn(o.shape, generator=gg) for o in outs]
    g_x, g_p = torch.randn(x.shape, generator=gg), torch.randn(p.shape, generator=gg)
    t_outs, t_x, t_p, grads = pto.reference_train_step(mt, sf, img, g_outs, g_x, g_p)
    blob.update({"train/out/x": t_x.numpy(), "train/out/projector": t_p.numpy(), "train/gin/x": g_x.numpy(),
                 "train/gin/projector": g_p.numpy()})
    for i, (o, g_) in enumerate(zip(t_outs, g_outs)):
        blob[f"train/out/head{i}"], blob[f"train/gin/head{i}"] = o.numpy(), g_.numpy()
    for k, v in grads.items():
        blob["train/grad/" + k] = v.numpy()
    for k, v in pto.tail_keys(mt.state_dict()).items():
        if "running" in k or "num_batches" in k:
            blob["train/state/" + k] = v.numpy()
    np.savez_compressed(out_path, **blob)
    print(f"wrote {out_path} ({out_path.stat().st_size/1e3:.1f} kB)")


if __name__ == "__main__":
    main()
