"""Host-side time breakdown of the end-to-end CRD step (pinned host inputs -> CRDLoss -> backward -> loss.item()).
Prints, per step (median of 40): wall time, time until forward returned, until backward returned, in .item()."""
import sys, time, json, contextlib, statistics
import torch
sys.path.insert(0, '.')
import bench
import __graft_entry__ as ge

pkg = ge.load_package()
dev = torch.device('cuda:0')
c = bench.HEADLINE
torch.manual_seed(bench.SEED)
with contextlib.redirect_stdout(sys.stderr):
    crit = pkg.CRDLoss(bench.make_opt(c)).to(dev)
host = bench.synth_inputs(c, torch, pin=True)[:3]
rows = []
for it in range(50):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    f_s, f_t, y = [t.to(dev, non_blocking=True) for t in host]
    f_s.requires_grad_()
    crit.zero_grad(set_to_none=True)
    t1 = time.perf_counter()
    with contextlib.redirect_stdout(sys.stderr):
        loss = crit(f_s, f_t, y, None)
    t2 = time.perf_counter()
    loss.backward()
    t3 = time.perf_counter()
    v = loss.item()
    t4 = time.perf_counter()
    if it >= 10:
        rows.append(((t4 - t0) * 1e6, (t1 - t0) * 1e6, (t2 - t1) * 1e6, (t3 - t2) * 1e6, (t4 - t3) * 1e6))
med = [statistics.median(r[i] for r in rows) for i in range(5)]
print(json.dumps({"us_per_step": med[0], "h2d_and_zero_grad_us": med[1], "forward_call_us": med[2],
                  "backward_call_us": med[3], "item_wait_us": med[4]}))
