set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
python profiles/r2_shard_step_once.py 8 > /dev/null 2>&1; echo plain rc=$?
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/r2_launches_shard8.csv python profiles/r2_shard_step_once.py 8 > gpurun_out/ncu11a.log 2>&1; echo rc=$?
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'crd_score_kernel|crd_finalize|crd_shard_filter' --launch-skip 9 -c 3 -o gpurun_out/r2_shard8_full python profiles/r2_shard_step_once.py 8 > gpurun_out/ncu11b.log 2>&1; echo rc=$?
ls -la gpurun_out
