set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_sharded_gpu.py tests/test_crd_gpu.py -m gpu -x -q 2>&1 | tail -5
timeout 300 python profiles/r2_shard_variant_sweep.py > gpurun_out/r2_shard_variant_sweep.json 2> gpurun_out/r2_shard_variant_sweep.err; cat gpurun_out/r2_shard_variant_sweep.json | tr -d '\n ' ; tail -3 gpurun_out/r2_shard_variant_sweep.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file gpurun_out/r2_launches_shard8_v2.csv python profiles/r2_shard_step_once.py 8 > gpurun_out/ncu14a.log 2>&1; echo rc=$?
grep -E "filter|finalize|score|allgather" gpurun_out/r2_launches_shard8_v2.csv | tail -4 | cut -d, -f5,15 
