"""Experiment: does the scoring pass get its bank rows from L2 when every anchor's negatives are row-sorted and the warp
partition is anchor-aligned?  Times crdpn_crd_step (event-bracketed kernel) for {unsorted, sorted} x {flat, aligned}."""
import ctypes, json, os, sys
import torch
QUICK = bool(os.environ.get("QUICK"))
sys.path.insert(0, '.')
import bench
import __graft_entry__ as ge

pkg = ge.load_package()
dev = torch.device('cuda:0')
lib = pkg._native.lib()
res = {}
for cname, c in ((("headline", bench.HEADLINE),) if QUICK else (("headline", bench.HEADLINE), ("config0", bench.CONFIG0))):
    torch.manual_seed(bench.SEED)
    crit = pkg.CRDLoss(bench.make_opt(c)).to(dev)
    f_s, f_t, y, cidx = [t.to(dev) for t in bench.synth_inputs(c, torch)]
    with torch.no_grad():
        v1 = crit.embed_s(f_s).contiguous(); v2 = crit.embed_t(f_t).contiguous()
    mem = crit.contrast
    mem._freeze_z(v1, v2, cidx)
    hp = mem._host_params()
    srt = cidx.clone()
    srt[:, 1:] = torch.sort(cidx[:, 1:], dim=1).values
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    for name, idx, var in (("unsorted_flat", cidx, 0), ("unsorted_aligned", cidx, 0x100), ("sorted_flat", srt, 0), ("sorted_aligned", srt, 0x100)):
        mem.variant = var
        for _ in range(1 if QUICK else 3):
            r = mem._step(v1, v2, y, idx, hp.Z1, hp.Z2)
        torch.cuda.synchronize()
        loss = r[0][5].item()
        lib.crdpn_timing_enable(1)
        tot, n = ctypes.c_double(), ctypes.c_uint64()
        lib.crdpn_timing_read(0, ctypes.byref(tot), ctypes.byref(n))
        for _ in range(2 if QUICK else 20):
            if cname == "config0":
                flush.fill_(1)
            mem._step(v1, v2, y, idx, hp.Z1, hp.Z2)
        torch.cuda.synchronize()
        lib.crdpn_timing_read(0, ctypes.byref(tot), ctypes.byref(n))
        lib.crdpn_timing_enable(0)
        res[f"{cname}/{name}"] = {"kernel_ms": tot.value / n.value, "loss": loss}
print(json.dumps(res, indent=1))
