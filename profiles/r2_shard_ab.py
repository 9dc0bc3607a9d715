"""Per-rank cost of the strong-scaling step on ONE GPU (rows [0, N/R) resident, whole replicated index list scanned):
score kernel and whole step (score + reduction/update) with and without the CTA-level fold of the warp partials
(variant bit 0x80 = one slot per warp and anchor, as in round 1)."""
import ctypes, json, sys
from pathlib import Path
import torch
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import __graft_entry__ as ge
from bench import HEADLINE, SEED

pkg = ge.load_package()
dev = torch.device("cuda:0")
lib = pkg._native.lib()
c = HEADLINE
g = torch.Generator().manual_seed(SEED)
v1 = torch.nn.functional.normalize(torch.randn(c["B"], c["D"], generator=g)).to(dev)
v2 = torch.nn.functional.normalize(torch.randn(c["B"], c["D"], generator=g)).to(dev)
y = torch.randperm(c["N"], generator=g)[:c["B"]].to(dev)
cidx = torch.randint(0, c["N"], (c["B"], c["K"] + 1), generator=g).to(dev)
cidx[:, 0] = y
cidx32 = cidx.to(torch.int32)
out = {}
for R in (1, 2, 4, 8):
    rows = c["N"] // R
    mem = pkg.ContrastMemory(c["D"], c["N"], c["K"], c["T"], c["m"], row_begin=0, row_end=rows).to(dev)
    for name, variant in (("fold", 0), ("no_fold", 0x80)):
        mem.variant = variant
        step = lambda: mem._step(v1, v2, y, cidx, 2.0e6, 2.0e6)
        for _ in range(5):
            step()
        torch.cuda.synchronize()
        tot, n = ctypes.c_double(), ctypes.c_uint64()
        lib.crdpn_timing_enable(1)
        lib.crdpn_timing_read(0, ctypes.byref(tot), ctypes.byref(n))
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(50):
            step()
        e1.record()
        torch.cuda.synchronize()
        lib.crdpn_timing_read(0, ctypes.byref(tot), ctypes.byref(n))
        lib.crdpn_timing_enable(0)
        # graph-captured step: what the multi-GPU bench replays (no host launch gaps)
        gph = torch.cuda.CUDAGraph()
        s = torch.cuda.Stream()
        s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s):
            step()
        torch.cuda.current_stream().wait_stream(s)
        with torch.cuda.graph(gph):
            step()
        for _ in range(5):
            gph.replay()
        torch.cuda.synchronize()
        g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        g0.record()
        for _ in range(50):
            gph.replay()
        g1.record()
        torch.cuda.synchronize()
        out[f"R{R}_{name}"] = {"step_ms": e0.elapsed_time(e1) / 50, "score_kernel_ms": tot.value / max(n.value, 1),
                               "graph_step_ms": g0.elapsed_time(g1) / 50}
        del gph
    del mem
    torch.cuda.empty_cache()
print(json.dumps(out, indent=1))
