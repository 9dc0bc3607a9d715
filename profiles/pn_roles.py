"""Per-role wait-cycle breakdown of pointnet_fwd_eval_kernel (variant bit2 = instrumentation on)."""
import sys, json, torch
sys.path.insert(0, '.')
import __graft_entry__ as ge
from oracle import pointnet_oracle as po
pkg = ge.load_package()
dev = torch.device('cuda:0')
st = po.random_state(1024, seed=46)
enc = pkg.ShapeEncoderPC(1024); enc.load_state_dict(st); enc = enc.to(dev).eval()
x = po.random_clouds(160, 2500, seed=46).to(dev)
enc.variant = 4
for _ in range(3): enc(x)
torch.cuda.synchronize()
ws = enc._ws
off = 160 * 1024 * 4
dbg = ws[off:off + 148 * 32 * 8].view(torch.int64).view(148, 32).cpu().double()
names = {0: "mma wait H1_FULL", 1: "mma wait A2_EMPTY", 2: "mma wait W3_FULL", 3: "mma wait H2_FULL", 4: "mma wait A3_EMPTY",
         5: "mma role total", 6: "producer wait W3_EMPTY", 8: "epiA wait A3_FULL", 9: "epiA total", 10: "epiB wait A3_FULL",
         11: "epiB total", 12: "front wait H1_EMPTY", 13: "front wait A2_FULL", 14: "front wait H2_EMPTY", 15: "front total",
         16: "front layer1 compute", 17: "front E2 compute", 18: "units"}
for k, n in names.items():
    col = dbg[:, k]
    print(f"{n:28s} mean {col.mean().item():12.0f}  min {col.min().item():12.0f}  max {col.max().item():12.0f}")
