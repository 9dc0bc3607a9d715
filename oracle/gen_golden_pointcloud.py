"""oracle/gen_golden_pointcloud.py -- make tests/golden/pointcloud_golden.npz from the REFERENCE's own read_pointcloud
(auxiliary/dataset.py:121-150; pymesh stubbed to hand back the seeded synthetic vertices, numpy's RNG seeded so the drawn
subset is known).  Build container only.  TEST INFRASTRUCTURE ONLY."""
from __future__ import annotations

import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from oracle import pointcloud_oracle as pco  # noqa: E402


def main(out_path: Path = ROOT / "tests" / "golden" / "pointcloud_golden.npz") -> None:
    meshes = pco.synthetic_meshes()
    blob = {"counts": np.array([m.shape[0] for m in meshes]), "point_num": np.int64(2500)}
    cases = [(0, 0.0, 11), (1, 37.0, 12), (2, 0.0, 13), (2, 215.0, 14), (0, 90.0, 15)]
    blob["cases"] = np.array(cases, dtype=np.float64)
    for j, (cid, rot, np_seed) in enumerate(cases):
        r = pco.call_reference(meshes[cid], 2500, rot, np_seed)
        if r is None:
            raise SystemExit("/root/reference is not mounted; golden vectors can only be made in the build container")
        blob[f"cloud{j}"], blob[f"subset{j}"] = r[0].numpy(), r[1]
    np.savez_compressed(out_path, **blob)
    print(f"wrote {out_path} ({out_path.stat().st_size/1e3:.1f} kB)")


if __name__ == "__main__":
    main()
