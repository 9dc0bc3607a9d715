import ctypes, json, sys, torch
sys.path.insert(0, '.')
import __graft_entry__ as ge
from bench import HEADLINE, SEED
pkg = ge.load_package(); dev = torch.device('cuda:0'); lib = pkg._native.lib(); c = HEADLINE
g = torch.Generator().manual_seed(SEED)
out = {}
for B in (46, 138):
    v1 = torch.nn.functional.normalize(torch.randn(B, c['D'], generator=g)).to(dev)
    v2 = torch.nn.functional.normalize(torch.randn(B, c['D'], generator=g)).to(dev)
    y = torch.randperm(c['N'], generator=g)[:B].to(dev)
    cidx = torch.randint(0, c['N'], (B, c['K'] + 1), generator=g).to(dev); cidx[:, 0] = y
    for name, variant, sweep in (('u8_plain', 0, False), ('u8_swept', 0, True), ('u4_swept', 4, True), ('u4_plain', 4, False)):
        mem = pkg.ContrastMemory(c['D'], c['N'], c['K'], c['T'], c['m'], bank_dtype=torch.bfloat16).to(dev)
        with torch.no_grad(): mem.params[2], mem.params[3] = 2.0e6, 2.0e6
        mem._host = None; mem.variant = variant; mem.streaming = False; mem.sweep = sweep
        step = lambda: mem._step(v1, v2, y, cidx, 2.0e6, 2.0e6)
        r = step(); torch.cuda.synchronize()
        for _ in range(5): step()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(30): step()
        e1.record(); torch.cuda.synchronize()
        out[f'B{B}_{name}'] = {'ms': round(e0.elapsed_time(e1) / 30, 4), 'loss': float(r[0][5])}
        del mem; torch.cuda.empty_cache()
print(json.dumps(out, indent=1))
