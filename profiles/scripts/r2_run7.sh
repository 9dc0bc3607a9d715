set -x
cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests/test_sharded_gpu.py tests/test_crd_gpu.py -m gpu -x -q 2>&1 | tail -30 > gpurun_out/r2_pytest7.log
tail -15 gpurun_out/r2_pytest7.log
