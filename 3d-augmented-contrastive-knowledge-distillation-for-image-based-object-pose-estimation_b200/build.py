"""In-tree build of libcrdpn_b200.so (sm_100a only) with plain nvcc.

The library is a C-ABI shared object (include/crdpn_b200.h); it does not link against torch.
Run:  python <package>/build.py   (or __graft_entry__.build()).
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys
from pathlib import Path

PKG = Path(__file__).resolve().parent
CSRC = PKG / "csrc"
LIB = PKG / "libcrdpn_b200.so"
SOURCES = ["host.cu", "crd_kernels.cu", "pointnet_kernels.cu", "pointnet_train.cu", "pointnet_backward.cu", "embed_kernels.cu", "p2p_kernels.cu", "crd_loss.cu", "kd_losses.cu", "pointcloud_sampler.cu", "crd_unfused.cu", "pointnet_train_split.cu", "pose_tail.cu"]
# development probes (include/crdpn_b200_dev.h): their own library, never linked into the product .so
DEV_LIB = PKG / "libcrdpn_b200_dev.so"
DEV_SOURCES = ["umma_tf32_probe.cu", "host.cu"]

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-std=c++17", "-lineinfo",
    "-Xcompiler", "-fPIC,-O2,-Wall",
    "--expt-relaxed-constexpr",
    "-Xptxas", "-v",
]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and Path(cand).exists():
            return cand
    raise RuntimeError("nvcc not found: libcrdpn_b200.so cannot be built (there is no CPU fallback)")


def sources() -> list[Path]:
    return [CSRC / s for s in SOURCES if (CSRC / s).exists()]


def needs_build() -> bool:
    if not LIB.exists() or not DEV_LIB.exists():
        return True
    t = min(LIB.stat().st_mtime, DEV_LIB.stat().st_mtime)
    deps = list(CSRC.glob("*.cu")) + list(CSRC.glob("*.cuh")) + [PKG.parent / "include" / "crdpn_b200.h"]
    return any(d.stat().st_mtime > t for d in deps)


def build(force: bool = False, verbose: bool = False) -> Path:
    if not force and not needs_build():
        return LIB
    objdir = PKG / "build"
    objdir.mkdir(exist_ok=True)
    nvcc = _nvcc()
    objs = []
    procs = []
    names = {s.name for s in sources()}
    for src in sources() + [CSRC / s for s in DEV_SOURCES if s not in names]:
        obj = objdir / (src.stem + ".o")
        cmd = [nvcc, *NVCC_FLAGS, "-c", str(src), "-o", str(obj)]
        procs.append((src, obj, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    log = []
    for src, obj, pr in procs:
        out, _ = pr.communicate()
        log.append(f"==== {src.name}\n{out}")
        if pr.returncode != 0:
            sys.stderr.write(out)
            raise RuntimeError(f"nvcc failed on {src.name}")
        objs.append(str(obj))
    (objdir / "ptxas.log").write_text("\n".join(log))
    if verbose:
        print("\n".join(log))
    dev_objs = [str(objdir / (Path(s).stem + ".o")) for s in DEV_SOURCES]
    prod_objs = [o for o in objs if Path(o).stem + ".cu" in names]
    arch = NVCC_FLAGS[:2]   # the link step too: without it nvcc adds an (empty) device-link stub for its default sm_52
    subprocess.check_call([nvcc, *arch, "-shared", "-o", str(LIB), *prod_objs, "-lcudart"])
    subprocess.check_call([nvcc, *arch, "-shared", "-o", str(DEV_LIB), *dev_objs, "-lcudart"])
    return LIB


if __name__ == "__main__":
    p = build(force="--force" in sys.argv, verbose="-v" in sys.argv)
    print(p)
