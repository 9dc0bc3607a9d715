"""One rank's share of the strong-scaling step (configs[3] over R shards) on ONE GPU, compact mode (filter pre-pass +
balanced ranges, variant 0x40) with every compiled shape of the scoring kernel (variant bits 0-2).  Prints JSON."""
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
out = {}
for R in (8, 4):
    rows = c['N'] // R
    for name, variant in [(f'v{v}', 0x40 | v) for v in range(8)]:
        m = pkg.ShardedContrastMemory(c['D'], c['N'], c['K'], rank=0, world_size=1, comm='p2p', seed=5).to(dev)
        m.row_begin, m.row_end = 0, rows
        m.memory_v1 = torch.nn.functional.normalize(torch.randn(rows, 128)).to(dev)
        m.memory_v2 = torch.nn.functional.normalize(torch.randn(rows, 128)).to(dev)
        m._relayout()
        with torch.no_grad(): m.params[2], m.params[3] = 2.0e6, 2.0e6
        m._host = None; m.variant = variant; m.fixed_local_batch = True
        o = m.step_resident(v1, v2, y, cidx)
        step = lambda: m.step_resident(v1, v2, y, cidx, o)
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
        out[f'R{R}_{name}'] = {'score_kernel_ms': round(tot.value / max(n.value, 1), 4), 'graph_step_ms': round(e0.elapsed_time(e1) / 50, 4)}
        del gph, m
        torch.cuda.empty_cache()
print(json.dumps(out, indent=1))
