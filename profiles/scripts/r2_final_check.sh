set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
python -c "import __graft_entry__ as ge; ge.smoke()" 2>&1 | grep -E "smoke|Error|error"
/usr/bin/time -v python bench.py > gpurun_out/r2_bench_final_default.json 2> gpurun_out/r2_bench_final_default.err; echo bench rc=$?
grep -E "Elapsed|Maximum resident" gpurun_out/r2_bench_final_default.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2_bench_final_default.json'))
print({k: d[k] for k in ('metric','value','unit','n_gpus','steps','warmup','ms_per_step','scaling','dtype','gpu_launches')})
print('roofline', d['roofline']); print('e2e', {k: d['e2e'][k] for k in ('value','ms_per_step','h2d_bytes_per_step','d2h_bytes_per_step')}); print('cpu', d['cpu_baseline']); print('clocks', d['clocks'])
PY
python bench.py --impl reference --steps 3 --warmup 1 | head -c 400
