"""Program profiled for the end-to-end CRD step through the public API (pinned host inputs -> CRDLoss -> backward)."""
import sys, torch
sys.path.insert(0, '.')
import bench
import __graft_entry__ as ge
pkg = ge.load_package()
dev = torch.device('cuda:0')
r = bench.time_crd_e2e(pkg, torch, dev, bench.HEADLINE, 5, 3)
print(r)
