set -x
cd $GRAFT_REPO_ROOT
python profiles/r2_shard_variants.py > gpurun_out/r2_shard_variants.json 2> gpurun_out/r2_shard_variants.err; cat gpurun_out/r2_shard_variants.json; tail -3 gpurun_out/r2_shard_variants.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:crd_score_kernel --launch-skip 3 -c 1 -o gpurun_out/r2_score_R8 python profiles/r2_shard_variants.py 8 > gpurun_out/ncu_r8.log 2>&1; echo rc=$?
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file gpurun_out/r2_launches_shard8.csv python profiles/r2_shard_variants.py 8 > /dev/null 2>&1; echo rc=$?
