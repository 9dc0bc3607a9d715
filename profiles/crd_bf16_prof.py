"""Program profiled for the bf16-bank scoring kernel at the headline config: 2 steps."""
import sys, torch
sys.path.insert(0, '.')
import bench
import __graft_entry__ as ge
pkg = ge.load_package()
dev = torch.device('cuda:0')
c = bench.HEADLINE
torch.manual_seed(bench.SEED)
crit = pkg.CRDLoss(bench.make_opt(c), bank_dtype=torch.bfloat16).to(dev)
f_s, f_t, y, cidx = [t.to(dev) for t in bench.synth_inputs(c, torch)]
with torch.no_grad():
    v1 = crit.embed_s(f_s).contiguous(); v2 = crit.embed_t(f_t).contiguous()
mem = crit.contrast
mem._freeze_z(v1, v2, cidx)
hp = mem._host_params()
for _ in range(2):
    mem._step(v1, v2, y, cidx, hp.Z1, hp.Z2)
torch.cuda.synchronize()
print("ok")
