set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_crd_gpu.py -m gpu -x -q 2>&1 | tail -25
timeout 700 python bench.py --steps 20 --warmup 5 > gpurun_out/r2_bench15.json 2> gpurun_out/r2_bench15.err; echo bench rc=$?
grep -v normalization gpurun_out/r2_bench15.err | tail -5
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2_bench15.json'))
print('value', d['value']/1e9, d['ms_per_step'])
e=d['e2e']; print('e2e graph', e['value']/1e9, e['ms_per_step'], 'no-graph', e['without_graph']['ms_per_step'], 'strict', e['sync_each_step']['ms_per_step'])
h=e['with_host_contrast_idx']; print('host idx graph', h['ms_per_step'], 'int32', h['int32_list']['ms_per_step'], 'nograph', h['without_graph']['ms_per_step'], 'strict', h['sync_each_step']['ms_per_step'])
print(json.dumps(d['also']['pose_tail'])[:900])
PY
