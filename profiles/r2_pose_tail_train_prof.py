import sys, time, torch
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
import __graft_entry__ as ge
from test_pose_tail_gpu import _random_reference_size_state
pkg = ge.load_package(); dev = torch.device('cuda:0')
sd = _random_reference_size_state(3)
tail = pkg.PoseTail(1024, 1024)
full = dict(sd)
for k in tail.state_dict():
    if k.endswith("num_batches_tracked"): full[k] = torch.zeros((), dtype=torch.long)
tail.load_state_dict(full); tail = tail.to(dev).train()
B = 138
sf = torch.randn(B, 1024, device=dev, requires_grad=True); img = torch.randn(B, 1024, device=dev, requires_grad=True)
def ev(): 
    e = torch.cuda.Event(enable_timing=True); e.record(); return e
for it in range(4):
    torch.cuda.synchronize(); t0 = time.perf_counter(); e0 = ev()
    outs, x, p = tail(sf, img)
    e1 = ev(); t1 = time.perf_counter()
    loss = sum(o.sum() for o in outs) + x.sum() + p.sum()
    e2 = ev()
    loss.backward()
    e3 = ev(); torch.cuda.synchronize(); t3 = time.perf_counter()
    print(f"fwd dev {e0.elapsed_time(e1)*1e3:.0f} us host {(t1-t0)*1e6:.0f} us | loss {e1.elapsed_time(e2)*1e3:.0f} | bwd dev {e2.elapsed_time(e3)*1e3:.0f} us | total wall {(t3-t0)*1e6:.0f} us")
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    outs, x, p = tail(sf, img); (sum(o.sum() for o in outs) + x.sum() + p.sum()).backward(); torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="self_cpu_time_total", row_limit=22, max_name_column_width=60))
