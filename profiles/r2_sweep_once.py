"""A few swept steps (variant | 0x400) at the headline shape: target of the ncu captures."""
import sys, torch
sys.path.insert(0, '.')
import __graft_entry__ as ge
from bench import HEADLINE, SEED
pkg = ge.load_package(); dev = torch.device('cuda:0'); c = HEADLINE
variant = int(sys.argv[1], 0) if len(sys.argv) > 1 else 0x400
g = torch.Generator().manual_seed(SEED)
v1 = torch.nn.functional.normalize(torch.randn(c['B'], c['D'], generator=g)).to(dev)
v2 = torch.nn.functional.normalize(torch.randn(c['B'], c['D'], generator=g)).to(dev)
y = torch.randperm(c['N'], generator=g)[:c['B']].to(dev)
cidx = torch.randint(0, c['N'], (c['B'], c['K'] + 1), generator=g).to(dev); cidx[:, 0] = y
mem = pkg.ContrastMemory(c['D'], c['N'], c['K'], c['T'], c['m']).to(dev)
with torch.no_grad(): mem.params[2], mem.params[3] = 2.0e6, 2.0e6
mem._host = None; mem.variant = variant
for _ in range(6): mem._step(v1, v2, y, cidx, 2.0e6, 2.0e6)
torch.cuda.synchronize()
