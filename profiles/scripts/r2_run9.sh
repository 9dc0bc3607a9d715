set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
nvidia-smi -L
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -30 > gpurun_out/r2_pytest9.log
tail -12 gpurun_out/r2_pytest9.log
timeout 700 python bench.py --steps 20 --warmup 5 > gpurun_out/r2_bench9.json 2> gpurun_out/r2_bench9.err; echo bench rc=$?
tail -c 600 gpurun_out/r2_bench9.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2_bench9_ref.json 2> gpurun_out/r2_bench9_ref.err; echo ref rc=$?
tail -c 300 gpurun_out/r2_bench9_ref.err
