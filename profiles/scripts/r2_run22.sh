set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
CRDPN_BENCH_QUICK=1 python bench.py --steps 3 --warmup 3 > /dev/null 2>&1; echo plain rc=$?
CRDPN_BENCH_QUICK=1 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/r2_launches_bench_quick.csv python bench.py --steps 3 --warmup 3 > gpurun_out/ncu22a.log 2>&1; echo rc=$?
CRDPN_BENCH_QUICK=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:crd_score_kernel --launch-skip 4 -c 1 -o gpurun_out/r2_score_full python bench.py --steps 3 --warmup 3 > gpurun_out/ncu22b.log 2>&1; echo rc=$?
python profiles/r2_pose_tail_prof.py > /dev/null 2>&1; echo plain2 rc=$?
timeout 600 ncu --set full --clock-control none --import-source on -k regex:pose_tail_kernel --launch-skip 6 -c 1 -o gpurun_out/r2_pose_tail_full python profiles/r2_pose_tail_prof.py > gpurun_out/ncu22c.log 2>&1; echo rc=$?
ls -la gpurun_out | tail -8
