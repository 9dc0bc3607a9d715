"""GraphedStep over the PoseTail train step exactly as bench_kd_losses.bench_pose_tail builds it, with the first exception
inside the capture printed (torch.cuda.graph.__exit__ replaces it by 'operation failed due to a previous error')."""
import sys, traceback, torch
sys.path.insert(0, '.')
import __graft_entry__ as ge
pkg = ge.load_package(); dev = torch.device('cuda:0')
B, Fs, Fi = 138, 1024, 1024
variant = sys.argv[1] if len(sys.argv) > 1 else "bench"
ptail = pkg.PoseTail(img_feature_dim=Fi, shape_feature_dim=Fs).to(dev)
ptail.train()
sf, img = torch.randn(B, Fs, device=dev), torch.randn(B, Fi, device=dev)
if variant == "bench":   # the eager run that precedes the capture in the bench
    sfg, imgg = sf.clone().requires_grad_(True), img.clone().requires_grad_(True)
    for _ in range(5):
        outs, x, p = ptail(sfg, imgg)
        (sum(o.sum() for o in outs) + x.sum() + p.sum()).backward()
    torch.cuda.synchronize()
first = []


def fwd_bwd(a, b):
    try:
        outs, x, p = ptail(a, b)
        loss = sum(o.sum() for o in outs) + x.sum() + p.sum()
        loss.backward()
        return loss
    except BaseException:
        first.append(traceback.format_exc())
        raise


try:
    host_in = (sf.cpu().pin_memory(), img.cpu().pin_memory())
    gs = pkg.GraphedStep(fwd_bwd, host_in, dev, grad_inputs=(0, 1), zero_grad=lambda: ptail.zero_grad(set_to_none=True))
    for _ in range(5):
        gs.stage(*host_in); gs.run(); gs.collect()
    torch.cuda.synchronize()
    import time
    t0 = time.perf_counter()
    for _ in range(50):
        gs.stage(*host_in); gs.run(); gs.collect()
    torch.cuda.synchronize()
    print(f"{variant}: GraphedStep OK, {(time.perf_counter() - t0) / 50 * 1e6:.0f} us per step", flush=True)
except BaseException as e:
    print(f"{variant}: FAILED: {type(e).__name__}: {str(e)[:300]}", flush=True)
    for t in first:
        print("first exception inside fn:\n" + t[-2500:], flush=True)
    if not first:
        traceback.print_exc()
