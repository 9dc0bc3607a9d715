"""Program profiled for the loss-code kernels and the point-cloud producer: 3 x (student step loss fwd+bwd, infoNCE_KD
fwd+bwd, poseNCE_KD fwd+bwd, one batch of 138 sampled clouds)."""
import sys
import numpy as np
import torch
sys.path.insert(0, '.')
import bench_kd_losses
import __graft_entry__ as ge

pkg = ge.load_package()
dev = torch.device('cuda:0')
out, tout, sf, tf, label = bench_kd_losses._synthetic_step(torch, 138, 200)
o = [t.to(dev).requires_grad_() for t in out]
to = [t.to(dev) for t in tout]
a, p, lab = sf.to(dev).requires_grad_(), tf.to(dev).requires_grad_(), label.to(dev)
rng = np.random.default_rng(0)
sampler = pkg.PointCloudSampler([rng.normal(size=(v, 3)) for v in (12000, 30000, 8000, 50000)], 2500, dev, seed=1)
ids = torch.randint(0, 4, (138,))
rots = torch.randint(0, 360, (138,)).float()
for _ in range(3):
    pkg.student_kd_step_loss(o, to, a, p, lab).backward()
    pkg.infoNCE_KD(a[:46], p[:46], None, 0.5).backward()
    pkg.poseNCE_KD(a[:46], p[:46], lab[:46], 0.5, "linear").backward()
    sampler.sample(ids, rots)
torch.cuda.synchronize()
print("ok")
