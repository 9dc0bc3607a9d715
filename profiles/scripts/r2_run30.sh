set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -6
timeout 700 python bench.py --steps 20 --warmup 5 > gpurun_out/r2_bench30.json 2> gpurun_out/r2_bench30.err; echo bench rc=$?
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2_bench30.json'))
print('value', d['value']/1e9, d['ms_per_step'], 'kernel', d['roofline']['kernel_ms'], 'frac', d['roofline']['frac'])
e=d['e2e']; print('e2e graph', e['value']/1e9, e['ms_per_step'], 'no-graph', e['without_graph']['ms_per_step'], 'strict', e['sync_each_step']['ms_per_step'])
h=e['with_host_contrast_idx']; print('host idx graph', h['ms_per_step'], 'int32', h['int32_list']['ms_per_step'])
a=d['also']; print('cfg0', a['config0_l2_flushed']['value']/1e9, a['config0_l2_warm']['value']/1e9, 'B138', a['B138_duplicate_idx']['value']/1e9, a['B138_duplicate_idx']['ms_per_step'])
print('bf16', a['headline_bf16_banks']['value']/1e9, a['headline_bf16_banks']['gather_kernel'])
print('shard emu', a['shard_emulation'])
PY
