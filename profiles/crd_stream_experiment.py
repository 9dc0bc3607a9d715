"""Gather vs bank-streaming formulation of the CRD step at the headline config and config 0: event-bracketed time of the
scoring part (for streaming: bucketing + streaming kernel), whole-step time, loss agreement."""
import ctypes, json, sys
import torch
sys.path.insert(0, '.')
import bench
import __graft_entry__ as ge

pkg = ge.load_package()
dev = torch.device('cuda:0')
lib = pkg._native.lib()
res = {}
for cname, c in (("headline", bench.HEADLINE), ("config0", bench.CONFIG0)):
    torch.manual_seed(bench.SEED)
    crit = pkg.CRDLoss(bench.make_opt(c)).to(dev)
    f_s, f_t, y, cidx = [t.to(dev) for t in bench.synth_inputs(c, torch)]
    with torch.no_grad():
        v1 = crit.embed_s(f_s).contiguous(); v2 = crit.embed_t(f_t).contiguous()
    mem = crit.contrast
    mem._freeze_z(v1, v2, cidx)
    hp = mem._host_params()
    bank0 = (mem.memory_v1.clone(), mem.memory_v2.clone())
    for name, streaming in (("gather", False), ("stream", True)):
        mem.streaming = streaming
        with torch.no_grad():
            mem.memory_v1.copy_(bank0[0]); mem.memory_v2.copy_(bank0[1])
        r = mem._step(v1, v2, y, cidx, hp.Z1, hp.Z2)
        torch.cuda.synchronize()
        loss = r[0][5].item()
        g1 = r[1].clone()
        for _ in range(3):
            mem._step(v1, v2, y, cidx, hp.Z1, hp.Z2)
        torch.cuda.synchronize()
        lib.crdpn_timing_enable(1)
        tot, n = ctypes.c_double(), ctypes.c_uint64()
        lib.crdpn_timing_read(0, ctypes.byref(tot), ctypes.byref(n))
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20):
            mem._step(v1, v2, y, cidx, hp.Z1, hp.Z2)
        e1.record()
        torch.cuda.synchronize()
        lib.crdpn_timing_read(0, ctypes.byref(tot), ctypes.byref(n))
        lib.crdpn_timing_enable(0)
        res[f"{cname}/{name}"] = {"score_part_ms": tot.value / n.value, "step_ms": e0.elapsed_time(e1) / 20, "first_loss": loss,
                                  "g1_absmax": g1.abs().max().item(), "g1_00": g1[0, 0].item()}
print(json.dumps(res, indent=1))
