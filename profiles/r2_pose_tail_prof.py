"""Pose-tail chain kernel at the KD-time shape (138 rows, 1024 + 1024 features): device time per call (CUDA events around a
batch of calls, and inside a CUDA graph), fp32-accurate and bf16 modes.  Also the target of the ncu captures."""
import json, sys, torch
sys.path.insert(0, '.')
import __graft_entry__ as ge
sys.path.insert(0, 'tests')
from test_pose_tail_gpu import _random_reference_size_state
pkg = ge.load_package(); dev = torch.device('cuda:0')
sd = _random_reference_size_state()
out = {}
for B in (138, 160, 46):
    sf, img = torch.randn(B, 1024, device=dev), torch.randn(B, 1024, device=dev)
    for name, dt in (('fp32_split', torch.float32), ('bf16', torch.bfloat16)):
        tail = pkg.FrozenPoseTail.from_state_dict(sd, dtype=dt).to(dev)
        for _ in range(5): tail(sf, img)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(100): tail(sf, img)
        e1.record(); torch.cuda.synchronize()
        eager_us = e0.elapsed_time(e1) * 10
        g = torch.cuda.CUDAGraph()
        s = torch.cuda.Stream(); s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s): tail(sf, img)
        torch.cuda.current_stream().wait_stream(s)
        with torch.cuda.graph(g):
            for _ in range(20): tail(sf, img)
        g.replay(); torch.cuda.synchronize()
        e0.record()
        for _ in range(10): g.replay()
        e1.record(); torch.cuda.synchronize()
        out[f'B{B}_{name}'] = {'per_call_us': round(eager_us, 2), 'in_graph_us': round(e0.elapsed_time(e1) * 1000 / 200, 2)}
print(json.dumps(out, indent=1))

# per-task timeline of one call (CRDPN_POSE_TAIL_PROF stamps), B = 138, fp32-accurate mode
import numpy as np
tail = pkg.FrozenPoseTail.from_state_dict(sd).to(dev)
sf, img = torch.randn(138, 1024, device=dev), torch.randn(138, 1024, device=dev)
for _ in range(3): tail(sf, img)
ch = tail._chain
torch.cuda.synchronize()
arr, outs = tail._outs[138]
ch.run(arr, sf, img, 4)
torch.cuda.synchronize()
ws = ch._ws[138]
stamps = ws[1024:1024 + 768 * 64].view(torch.int64).cpu().numpy().reshape(768, 8)
stamps = stamps[stamps[:, 0] > 0]
t0 = stamps.min()
rel = (stamps - t0) / 1000.0
print("tasks", len(rel), file=sys.stderr)
names = ["acc_ready", "tmem_drained", "stored", "tile_complete", "reduce_left", "staged_all", "staged_own", "done"]
bounds = [0, 140, 248, 326, 342, len(rel)]   # dependency levels are contiguous in task order
for a, b in zip(bounds[:-1], bounds[1:]):
    r = rel[a:b]
    if len(r):
        print(f"tasks {a}-{b}: " + "  ".join(f"{nm} {r[:, k].min():.1f}..{r[:, k].max():.1f}" for k, nm in enumerate(names)), file=sys.stderr)
        order = [0, 1, 2, 3, 6, 5, 4, 7]   # time order of the stamps
        d = np.diff(r[:, order], axis=1)
        print("    mean step durations (us): " + "  ".join(f"{names[order[k + 1]]} {d[:, k].mean():.2f}" for k in range(7)), file=sys.stderr)
