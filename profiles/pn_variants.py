import sys, json, torch, ctypes
sys.path.insert(0, '.')
import __graft_entry__ as ge
from oracle import pointnet_oracle as po
pkg = ge.load_package(); lib = pkg._native.lib()
dev = torch.device('cuda:0')
st = po.random_state(1024, seed=46)
enc = pkg.ShapeEncoderPC(1024); enc.load_state_dict(st); enc = enc.to(dev).eval()
x = po.random_clouds(160, 2500, seed=46).to(dev)
ref = None
for variant in (0, 8):
    enc.variant = variant
    for _ in range(5): out = enc(x)
    torch.cuda.synchronize()
    lib.crdpn_timing_enable(1)
    tot, n = ctypes.c_double(), ctypes.c_uint64()
    lib.crdpn_timing_read(1, ctypes.byref(tot), ctypes.byref(n))
    for _ in range(50): out = enc(x)
    torch.cuda.synchronize()
    lib.crdpn_timing_read(1, ctypes.byref(tot), ctypes.byref(n))
    lib.crdpn_timing_enable(0)
    if ref is None: ref = out.clone()
    print(json.dumps({"variant": variant, "kernel_ms": tot.value / n.value, "maxdiff_vs_v0": (out - ref).abs().max().item()}))
