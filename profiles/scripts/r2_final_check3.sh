cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
SECONDS=0
timeout 400 python -m pytest tests -m gpu -x -q 2>&1 | tail -4 | tee gpurun_out/r2_pytest_gpu_final.log
echo pytest elapsed ${SECONDS}s
python -c "import __graft_entry__ as ge; ge.smoke()" 2>&1 | grep -E "smoke|Error|error"
echo smoke elapsed ${SECONDS}s
timeout 300 python bench.py > gpurun_out/r2_bench_final_rebuilt.json 2> gpurun_out/r2_bench_final_rebuilt.err; echo bench rc=$? elapsed ${SECONDS}s
head -c 700 gpurun_out/r2_bench_final_rebuilt.json
