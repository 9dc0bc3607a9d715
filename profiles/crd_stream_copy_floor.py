"""How fast can the bank be streamed through the shared-memory ring at all (variant 0xA00: tiles bulk-copied, nothing
scored)?  The floor of the bank-streaming formulation with this ring (1 CTA / SM, 4 x 36 KB stages)."""
import ctypes, json, sys
import torch
sys.path.insert(0, '.')
import bench
import __graft_entry__ as ge
pkg = ge.load_package()
dev = torch.device('cuda:0')
lib = pkg._native.lib()
c = bench.HEADLINE
torch.manual_seed(bench.SEED)
crit = pkg.CRDLoss(bench.make_opt(c)).to(dev)
f_s, f_t, y, cidx = [t.to(dev) for t in bench.synth_inputs(c, torch)]
with torch.no_grad():
    v1 = crit.embed_s(f_s).contiguous(); v2 = crit.embed_t(f_t).contiguous()
mem = crit.contrast
mem._freeze_z(v1, v2, cidx)
hp = mem._host_params()
res = {}
for name, var in (("stream_full", 0x200), ("stream_copy_only", 0xA00)):
    mem.variant = var
    for _ in range(2):
        mem._step(v1, v2, y, cidx, hp.Z1, hp.Z2)
    torch.cuda.synchronize()
    lib.crdpn_timing_enable(1)
    tot, n = ctypes.c_double(), ctypes.c_uint64()
    lib.crdpn_timing_read(0, ctypes.byref(tot), ctypes.byref(n))
    for _ in range(10):
        mem._step(v1, v2, y, cidx, hp.Z1, hp.Z2)
    torch.cuda.synchronize()
    lib.crdpn_timing_read(0, ctypes.byref(tot), ctypes.byref(n))
    lib.crdpn_timing_enable(0)
    res[name] = {"bucketing_plus_stream_ms": tot.value / n.value}
print(json.dumps(res))
