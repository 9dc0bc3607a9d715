set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
nvidia-smi -L
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -40 > gpurun_out/r2_pytest1.log
cat gpurun_out/r2_pytest1.log | tail -30
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r2_bench1.json 2> gpurun_out/r2_bench1.err; echo bench rc=$?
tail -c 1500 gpurun_out/r2_bench1.err
CRDPN_FORCE_MULTI=1 timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r2_bench1_multi1.json 2> gpurun_out/r2_bench1_multi1.err; echo multi rc=$?
tail -c 2500 gpurun_out/r2_bench1_multi1.err
