"""Score-kernel variant sweep on one rank's share of the strong-scaling step (rows [0, N/R) resident)."""
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
only = int(sys.argv[1]) if len(sys.argv) > 1 else None
out = {}
for R in ((only,) if only else (8, 4, 1)):
    rows = c["N"] // R
    mem = pkg.ContrastMemory(c["D"], c["N"], c["K"], c["T"], c["m"], row_begin=0, row_end=rows).to(dev)
    for variant in ((0,) if only else range(0, 8)):
        mem.variant = variant
        step = lambda: mem._step(v1, v2, y, cidx, 2.0e6, 2.0e6)
        try:
            for _ in range(3):
                step()
        except RuntimeError as exc:
            out[f"R{R}_v{variant}"] = str(exc)[:60]
            continue
        torch.cuda.synchronize()
        tot, n = ctypes.c_double(), ctypes.c_uint64()
        lib.crdpn_timing_enable(1)
        lib.crdpn_timing_read(0, ctypes.byref(tot), ctypes.byref(n))
        for _ in range(30):
            step()
        torch.cuda.synchronize()
        lib.crdpn_timing_read(0, ctypes.byref(tot), ctypes.byref(n))
        lib.crdpn_timing_enable(0)
        out[f"R{R}_v{variant}"] = round(tot.value / max(n.value, 1), 4)
    del mem
    torch.cuda.empty_cache()
print(json.dumps(out, indent=1))
