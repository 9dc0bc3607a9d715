cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
SECONDS=0
timeout 300 python -m pytest tests/test_sharded_gpu.py tests/test_pointnet_sync_gpu.py -m gpu -x -q 2>&1 | tail -4 | tee gpurun_out/r2_pytest_2gpu_final.log
echo pytest elapsed ${SECONDS}s
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 200 --warmup 5 > gpurun_out/r2_bench_2gpu_final.json 2> gpurun_out/r2_bench_2gpu_final.err; echo bench2 rc=$? elapsed ${SECONDS}s
grep -v "normalization" gpurun_out/r2_bench_2gpu_final.err | tail -5
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2_bench_2gpu_final.json'))
print('value %.2f G %.4f ms scaling %s parity %s' % (d['value']/1e9, d['ms_per_step'], d['scaling'], d['parity']['ok']))
e=d['e2e']; print('e2e %.2f G %.4f ms' % (e['value']/1e9, e['ms_per_step']))
PY
