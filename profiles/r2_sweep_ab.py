"""Band-sorted ("swept") lists (variant | 0x400: crd_band_sort_kernel + interleaved blocks) against the default gather step:
headline shape on one GPU, and one rank's share of the strong-scaling step at R = 2 / 4 / 8.  Step time in a CUDA graph, the
scoring kernel's own time, and the agreement of loss / gradients / updated rows between the two modes.  Prints JSON."""
import ctypes, json, sys, torch
sys.path.insert(0, '.')
import __graft_entry__ as ge
from bench import HEADLINE, SEED
pkg = ge.load_package(); dev = torch.device('cuda:0'); lib = pkg._native.lib(); c = HEADLINE
g = torch.Generator().manual_seed(SEED)
v1 = torch.nn.functional.normalize(torch.randn(c['B'], c['D'], generator=g)).to(dev)
v2 = torch.nn.functional.normalize(torch.randn(c['B'], c['D'], generator=g)).to(dev)
y = torch.randperm(c['N'], generator=g)[:c['B']].to(dev)
cidx = torch.randint(0, c['N'], (c['B'], c['K'] + 1), generator=g).to(dev); cidx[:, 0] = y
bank = torch.nn.functional.normalize(torch.randn(c['N'], 2, 128, generator=g), dim=2)
out = {}

def timed(step):
    for _ in range(5): step()
    torch.cuda.synchronize()
    tot, n = ctypes.c_double(), ctypes.c_uint64()
    lib.crdpn_timing_enable(1); lib.crdpn_timing_read(0, ctypes.byref(tot), ctypes.byref(n))
    for _ in range(30): step()
    torch.cuda.synchronize()
    lib.crdpn_timing_read(0, ctypes.byref(tot), ctypes.byref(n)); lib.crdpn_timing_enable(0)
    gph = torch.cuda.CUDAGraph(); s = torch.cuda.Stream(); s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s): step()
    torch.cuda.current_stream().wait_stream(s)
    with torch.cuda.graph(gph): step()
    for _ in range(5): gph.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(50): gph.replay()
    e1.record(); torch.cuda.synchronize()
    return {'score_kernel_ms': round(tot.value / max(n.value, 1), 4), 'graph_step_ms': round(e0.elapsed_time(e1) / 50, 4)}

# ---- one GPU, whole bank
res = {}
for name, variant in [('gather', 0), ('swept', 0x400)]:
    mem = pkg.ContrastMemory(c['D'], c['N'], c['K'], c['T'], c['m']).to(dev)
    with torch.no_grad():
        mem.memory_v1.copy_(bank[:, 0]); mem.memory_v2.copy_(bank[:, 1]); mem.params[2], mem.params[3] = 2.0e6, 2.0e6
    mem._host = None; mem.variant = variant
    r = mem._step(v1, v2, y, cidx, 2.0e6, 2.0e6)
    torch.cuda.synchronize()
    res[name] = (r[0].clone(), r[1].clone(), r[2].clone(), mem.memory_v1[y].clone())
    with torch.no_grad():
        mem.memory_v1.copy_(bank[:, 0]); mem.memory_v2.copy_(bank[:, 1])
    out['1gpu_' + name] = timed(lambda: mem._step(v1, v2, y, cidx, 2.0e6, 2.0e6))
    del mem; torch.cuda.empty_cache()
a, b = res['gather'], res['swept']
rel = lambda x, z: ((x.double() - z.double()).abs().max() / z.double().abs().max()).item()
out['1gpu_agreement'] = {'loss_rel': abs((a[0][5] - b[0][5]).item()) / abs(a[0][5].item()), 'grad_v1_rel': rel(b[1], a[1]),
                         'grad_v2_rel': rel(b[2], a[2]), 'updated_rows_bit_identical': bool(torch.equal(a[3], b[3]))}
# ---- one rank's share at R shards
for R in (2, 4, 8):
    rows = c['N'] // R
    for name, variant in (('scan', 0x20), ('compact', 0x40), ('swept', 0x400)):
        m = pkg.ShardedContrastMemory(c['D'], c['N'], c['K'], rank=0, world_size=1, comm='p2p', seed=5).to(dev)
        m.row_begin, m.row_end = 0, rows
        m.memory_v1 = bank[:rows, 0].contiguous().to(dev); m.memory_v2 = bank[:rows, 1].contiguous().to(dev)
        m._relayout()
        with torch.no_grad(): m.params[2], m.params[3] = 2.0e6, 2.0e6
        m._host = None; m.variant = variant; m.fixed_local_batch = True
        o = m.step_resident(v1, v2, y, cidx)
        torch.cuda.synchronize()
        key = f'R{R}_{name}'
        res[key] = o['reduced'].clone()
        out[key] = timed(lambda: m.step_resident(v1, v2, y, cidx, o))
        del m; torch.cuda.empty_cache()
    out[f'R{R}_agreement'] = rel(res[f'R{R}_swept'][:2 * 46 * 128], res[f'R{R}_compact'][:2 * 46 * 128])
    mem_probe = None
print(json.dumps(out, indent=1))
