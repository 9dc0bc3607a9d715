set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
nvidia-smi -L
timeout 300 python profiles/r2_shard_compact.py > gpurun_out/r2_shard_compact.json 2> gpurun_out/r2_shard_compact.err; cat gpurun_out/r2_shard_compact.json; tail -3 gpurun_out/r2_shard_compact.err
timeout 900 python -m pytest tests/test_sharded_gpu.py tests/test_pointnet_sync_gpu.py -m gpu -x -q 2>&1 | tail -30 > gpurun_out/r2_pytest10_2gpu.log
tail -12 gpurun_out/r2_pytest10_2gpu.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r2_bench10_2gpu.json 2> gpurun_out/r2_bench10_2gpu.err; echo bench2 rc=$?
grep -v "normalization" gpurun_out/r2_bench10_2gpu.err | tail -15
head -c 4000 gpurun_out/r2_bench10_2gpu.json
