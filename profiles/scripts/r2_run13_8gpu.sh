set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
nvidia-smi -L | head -8
for N in 8 4; do
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2951$N bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/r2_bench13_${N}gpu.json 2> gpurun_out/r2_bench13_${N}gpu.err; echo bench$N rc=$?
grep -v "normalization" gpurun_out/r2_bench13_${N}gpu.err | tail -5
head -c 1500 gpurun_out/r2_bench13_${N}gpu.json
done
