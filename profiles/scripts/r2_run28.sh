set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_crd_gpu.py tests/test_sharded_gpu.py tests/test_guard_bands_gpu.py -m gpu -x -q 2>&1 | tail -8
CRDPN_BENCH_QUICK=1 python bench.py --steps 50 --warmup 5 2>/dev/null
python profiles/r2_shard_compact.py 2>/dev/null | tr -d '\n '
