import sys, torch
sys.path.insert(0, '.')
import __graft_entry__ as ge
from oracle import pointnet_oracle as po
pkg = ge.load_package()
dev = torch.device('cuda:0')
st = po.random_state(1024, seed=46)
enc = pkg.ShapeEncoderPC(1024); enc.load_state_dict(st); enc = enc.to(dev).eval()
x = po.random_clouds(160, 2500, seed=46).to(dev)
for _ in range(5): enc(x)
torch.cuda.synchronize()
print("ok")
