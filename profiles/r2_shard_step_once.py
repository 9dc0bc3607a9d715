"""One rank's share of the strong-scaling step at R shards (default 8) on one GPU, a few eager steps: target of the ncu
launch list / full capture (profiles/scripts/r2_run11.sh)."""
import sys, torch
sys.path.insert(0, '.')
import __graft_entry__ as ge
from bench import HEADLINE, SEED
R = int(sys.argv[1]) if len(sys.argv) > 1 else 8
variant = int(sys.argv[2], 0) if len(sys.argv) > 2 else 0
pkg = ge.load_package(); dev = torch.device('cuda:0'); c = HEADLINE
g = torch.Generator().manual_seed(SEED)
v1 = torch.nn.functional.normalize(torch.randn(c['B'], c['D'], generator=g)).to(dev)
v2 = torch.nn.functional.normalize(torch.randn(c['B'], c['D'], generator=g)).to(dev)
y = torch.randperm(c['N'], generator=g)[:c['B']].to(dev)
cidx = torch.randint(0, c['N'], (c['B'], c['K'] + 1), generator=g).to(dev); cidx[:, 0] = y
rows = c['N'] // R
m = pkg.ShardedContrastMemory(c['D'], c['N'], c['K'], rank=0, world_size=1, comm='p2p', seed=5).to(dev)
m.row_begin, m.row_end = 0, rows
m.memory_v1 = torch.nn.functional.normalize(torch.randn(rows, 128)).to(dev)
m.memory_v2 = torch.nn.functional.normalize(torch.randn(rows, 128)).to(dev)
m._relayout()
with torch.no_grad(): m.params[2], m.params[3] = 2.0e6, 2.0e6
m._host = None; m.variant = variant; m.fixed_local_batch = True
o = m.step_resident(v1, v2, y, cidx)
for _ in range(6): m.step_resident(v1, v2, y, cidx, o)
torch.cuda.synchronize()
