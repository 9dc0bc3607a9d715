set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r2_bench17_2gpu.json 2> gpurun_out/r2_bench17_2gpu.err; echo bench2 rc=$?
grep -v "normalization" gpurun_out/r2_bench17_2gpu.err | tail -12
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2_bench17_2gpu.json'))
print('value %.2f G %.4f ms parity %s' % (d['value']/1e9, d['ms_per_step'], d['parity']['ok']))
e=d['e2e']; print('e2e graph %.2f G %.4f ms | strict %.4f ms' % (e['value']/1e9, e['ms_per_step'], e['sync_each_step']['ms_per_step']))
h=e['with_host_contrast_idx']; print('host idx graph %.4f ms strict %.4f' % (h['ms_per_step'], h['sync_each_step']['ms_per_step']))
print(json.dumps(d['also']['pointnet'].get('train_sync_batchnorm'))[:600])
PY
