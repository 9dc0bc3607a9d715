cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
SECONDS=0
timeout 120 python -m pytest tests/test_pose_tail_gpu.py -m gpu -q -k "frees_its_activations" 2>&1 | tail -15 | cut -c1-300
echo newtest elapsed ${SECONDS}s
timeout 400 python -m pytest tests -m gpu -x -q 2>&1 | tail -4 | tee gpurun_out/r2_pytest_gpu_final5.log
echo pytest elapsed ${SECONDS}s
python -c "import __graft_entry__ as ge; ge.smoke()" 2>&1 | grep -E "smoke|Error|error"
timeout 300 python bench.py > gpurun_out/r2_bench_final5.json 2> gpurun_out/r2_bench_final5.err; echo bench rc=$? elapsed ${SECONDS}s
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2_bench_final5.json'))
print('value %.2f G %.4f ms e2e %.2f G' % (d['value']/1e9, d['ms_per_step'], d['e2e']['value']/1e9))
p=d['also']['pose_tail']; print({k: (v if not isinstance(v,str) else v[:80]) for k,v in p.items() if k.startswith('train')})
PY
