set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/r2_pytest20.log; cat gpurun_out/r2_pytest20.log
python -c "import __graft_entry__ as ge; ge.smoke()" 2>&1 | grep smoke
timeout 600 python bench_config3.py > gpurun_out/r2_config3.json 2> gpurun_out/r2_config3.err; echo config3 rc=$?; head -c 1500 gpurun_out/r2_config3.json; tail -3 gpurun_out/r2_config3.err
