"""Is kind::tf32's operand conversion a truncation (biased) or a rounding (unbiased)?  All-positive operands through the
probe: mean signed relative error of the 128 x 96 scores against fp64, and against two emulations of the conversion."""
import json, sys
import numpy as np
import torch
sys.path.insert(0, '.')
import __graft_entry__ as ge

pkg = ge.load_package()
lib = pkg._native.dev_lib()
dev = torch.device('cuda:0')
rng = np.random.default_rng(1)
rows1, rows2 = rng.uniform(0.5, 1.5, (64, 128)).astype(np.float32), rng.uniform(0.5, 1.5, (64, 128)).astype(np.float32)
v1, v2 = rng.uniform(0.5, 1.5, (48, 128)).astype(np.float32), rng.uniform(0.5, 1.5, (48, 128)).astype(np.float32)
z = np.zeros((64, 48), np.float32)
d = [torch.from_numpy(a).to(dev) for a in (rows1, rows2, v1, v2, z, z)]
out = torch.empty(128, 192, device=dev)
assert lib.crdpn_umma_tf32_probe(*[t.data_ptr() for t in d], out.data_ptr(), 0, torch.cuda.current_stream().cuda_stream) == 0
torch.cuda.synchronize()
got = out.cpu().numpy().astype(np.float64)[:, :96]
A = np.concatenate([rows1, rows2])
V = np.concatenate([v2, v1])


def trunc(x):
    return (x.view(np.uint32) & np.uint32(0xFFFFE000)).view(np.float32)


def rne(x):
    u = x.view(np.uint32).astype(np.uint64)
    u = (u + 0xFFF + ((u >> 13) & 1)) & 0xFFFFE000
    return u.astype(np.uint32).view(np.float32)


exact = A.astype(np.float64) @ V.astype(np.float64).T
res = {"mean_signed_rel_err_vs_fp64": float(((got - exact) / exact).mean()),
       "max_abs_rel_err_vs_fp64": float(np.abs((got - exact) / exact).max())}
for name, f in (("truncate", trunc), ("round_nearest_even", rne)):
    emu = f(A).astype(np.float64) @ f(V).astype(np.float64).T
    res[f"max_rel_diff_vs_{name}_emulation"] = float(np.abs((got - emu) / emu).max())
print(json.dumps(res, indent=1))
