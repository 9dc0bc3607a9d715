cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
SECONDS=0
python bench.py > gpurun_out/r2_bench_final_default.json 2> gpurun_out/r2_bench_final_default.err; echo bench rc=$? elapsed ${SECONDS}s
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2_bench_final_default.json'))
print({k: d[k] for k in ('metric','value','unit','n_gpus','steps','warmup','ms_per_step','scaling','dtype','gpu_launches')})
print('roofline', d['roofline']); print('e2e', {k: d['e2e'][k] for k in ('value','ms_per_step','h2d_bytes_per_step','d2h_bytes_per_step')}); print('cpu', d['cpu_baseline']); print('clocks', d['clocks'])
PY
