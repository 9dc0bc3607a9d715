"""Diagnostic: train-mode PointNet forward/backward on the GPU vs the oracle (fp64 autograd of the restated module).
Prints one error figure per quantity so one gpurun call localises a bug.   python profiles/pn_train_check.py [B P F]"""
import sys
import time

import torch

sys.path.insert(0, '.')
import __graft_entry__ as ge
from oracle import pointnet_oracle as po

pkg = ge.load_package()
dev = torch.device('cuda:0')


def nrel(got, want):
    got, want = got.double().cpu(), want.double().cpu()
    return ((got - want).norm() / (want.norm() + 1e-30)).item()


def run(B, P, F, seed=5, dtype=torch.float64):
    st = po.random_state(F, seed=seed)
    x = po.random_clouds(B, P, seed=seed + 1)
    gout = torch.randn(B, F, generator=torch.Generator().manual_seed(seed + 2))
    # oracle: autograd through the restated forward
    params = {k: v.clone().to(dtype).requires_grad_(v.is_floating_point() and 'running' not in k) for k, v in st.items()}
    new_stats = {}
    want = po.forward(x, params, training=True, new_stats=new_stats, dtype=dtype)
    (want * gout.to(dtype)).sum().backward()
    # same precision recipe as the kernels (bf16 operands, straight-through): separates arg-max flips from bugs
    eparams = {k: v.clone().to(dtype).requires_grad_(v.is_floating_point() and 'running' not in k) for k, v in st.items()}
    ewant = po.forward_train_bf16_emulated(x, eparams, dtype=dtype)
    (ewant * gout.to(dtype)).sum().backward()
    enc = pkg.ShapeEncoderPC(F)
    enc.load_state_dict({k: v.clone() for k, v in st.items()})
    enc = enc.to(dev).train()
    t0 = time.time()
    out = enc(x.to(dev))
    (out * gout.to(dev)).sum().backward()
    torch.cuda.synchronize()
    print(f"--- B={B} P={P} F={F}  ({time.time()-t0:.2f}s)")
    print(f"  out  vs emulated nrel {nrel(out.detach(), ewant.detach()):.3e}")
    print(f"  out            nrel {nrel(out.detach(), want.detach()):.3e}   maxrel {((out.detach().cpu().double()-want.detach()).abs().max()/want.detach().abs().max()).item():.3e}")
    for n in (1, 2, 3):
        bn = getattr(enc, f"bn{n}")
        print(f"  bn{n}.running_mean nrel {nrel(bn.running_mean, new_stats[f'bn{n}.running_mean']):.3e}  "
              f"running_var nrel {nrel(bn.running_var, new_stats[f'bn{n}.running_var']):.3e}  nbt {int(bn.num_batches_tracked)}")
    for name, prm in enc.named_parameters():
        w = params[name].grad
        g = prm.grad
        if 'conv' in name and 'bias' in name:
            print(f"  grad {name:14s} max|got| {g.abs().max().item():.3e}  (want ~0: {w.abs().max().item():.3e})")
        else:
            print(f"  grad {name:14s} vs fp oracle {nrel(g, w):.3e}   vs bf16-emulated oracle {nrel(g, eparams[name].grad):.3e}   |want| {w.norm().item():.3e}")


if __name__ == "__main__":
    if len(sys.argv) == 4:
        run(int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3]))
    else:
        for B, P, F in [(2, 50, 128), (3, 333, 1024), (4, 1000, 256), (16, 2500, 1024)]:
            run(B, P, F, dtype=torch.float64 if B * P * F < 5e6 else torch.float32)
