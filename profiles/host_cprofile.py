"""cProfile of the host side of (a) the end-to-end CRD step and (b) the fused student-step loss, 300 calls each."""
import cProfile, pstats, sys, io, contextlib
import torch
sys.path.insert(0, '.')
import bench, bench_kd_losses
import __graft_entry__ as ge

pkg = ge.load_package()
dev = torch.device('cuda:0')
c = bench.HEADLINE
with contextlib.redirect_stdout(sys.stderr):
    crit = pkg.CRDLoss(bench.make_opt(c)).to(dev)
host = bench.synth_inputs(c, torch, pin=True)[:3]


def crd_step():
    f_s, f_t, y = [t.to(dev, non_blocking=True) for t in host]
    f_s.requires_grad_()
    crit.zero_grad(set_to_none=True)
    loss = crit(f_s, f_t, y, None)
    loss.backward()
    return loss.item()


out, tout, sf, tf, label = bench_kd_losses._synthetic_step(torch, 138, 200)
o = [t.to(dev).requires_grad_() for t in out]
to = [t.to(dev) for t in tout]
a, p, lab = sf.to(dev).requires_grad_(), tf.to(dev), label.to(dev)


def mixer():
    pkg.student_kd_step_loss(o, to, a, p, lab).backward()


for name, fn in (("crd_step", crd_step), ("mixer", mixer)):
    with contextlib.redirect_stdout(sys.stderr):
        for _ in range(20):
            fn()
    torch.cuda.synchronize()
    pr = cProfile.Profile()
    pr.enable()
    for _ in range(300):
        fn()
    pr.disable()
    s = io.StringIO()
    pstats.Stats(pr, stream=s).sort_stats("tottime").print_stats(22)
    print(f"==== {name}\n" + s.getvalue()[:6000])
