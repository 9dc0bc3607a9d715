set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_crd_gpu.py tests/test_sharded_gpu.py tests/test_abi.py tests/test_crd_stream_gpu.py tests/test_guard_bands_gpu.py -m gpu -x -q 2>&1 | tail -30 > gpurun_out/r2_pytest2.log
tail -15 gpurun_out/r2_pytest2.log
timeout 300 python profiles/r2_shard_ab.py > gpurun_out/r2_shard_ab.json 2> gpurun_out/r2_shard_ab.err; echo ab rc=$?
cat gpurun_out/r2_shard_ab.json; tail -3 gpurun_out/r2_shard_ab.err
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r2_bench2.json 2> gpurun_out/r2_bench2.err; echo bench rc=$?
tail -c 600 gpurun_out/r2_bench2.err
# ncu: launch list of one fp32-accurate PointNet train step + full sets of its three big kernels
cat > /tmp/pn_train_once.py <<'PY'
import sys, torch
sys.path.insert(0, '.')
import __graft_entry__ as ge
from oracle import pointnet_oracle as po
pkg = ge.load_package()
dev = torch.device('cuda:0')
st = po.random_state(1024, seed=46)
enc = pkg.ShapeEncoderPC(1024); enc.load_state_dict(st); enc = enc.to(dev).train()
x = po.random_clouds(160, 2500, seed=46).to(dev)
g = torch.randn(160, 1024, device=dev)
for _ in range(3):
    for p in enc.parameters(): p.grad = None
    enc(x).backward(g)
torch.cuda.synchronize()
PY
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/r2_launches_pointnet_train_fp32.csv python /tmp/pn_train_once.py > gpurun_out/ncu_pn1.log 2>&1; echo ncu1 rc=$?
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'pn_bwd_pass2_kernel|pn_bwd_pass3_kernel|pointnet_fwd_train_split_kernel|pn_stats2_split_kernel' --launch-skip 8 -c 4 -o gpurun_out/r2_pn_train_full python /tmp/pn_train_once.py > gpurun_out/ncu_pn2.log 2>&1; echo ncu2 rc=$?
ls -la gpurun_out | tail -12
