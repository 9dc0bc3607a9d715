"""Program profiled for the train-mode PointNet step (B=160, P=2500, F=1024): 3 x (forward + backward)."""
import sys, torch
sys.path.insert(0, '.')
import __graft_entry__ as ge
from oracle import pointnet_oracle as po
pkg = ge.load_package()
dev = torch.device('cuda:0')
st = po.random_state(1024, seed=46)
enc = pkg.ShapeEncoderPC(1024); enc.load_state_dict(st); enc = enc.to(dev).train()
x = po.random_clouds(160, 2500, seed=46).to(dev)
gout = torch.randn(160, 1024, device=dev)
for _ in range(3):
    for p in enc.parameters(): p.grad = None
    enc(x).backward(gout)
torch.cuda.synchronize()
print("ok")
