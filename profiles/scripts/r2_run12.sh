set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_pose_tail_gpu.py -m gpu -x -q 2>&1 | tail -40 > gpurun_out/r2_pytest12.log
tail -40 gpurun_out/r2_pytest12.log
timeout 300 python profiles/r2_pose_tail_prof.py > gpurun_out/r2_pose_tail_prof.json 2> gpurun_out/r2_pose_tail_prof.err; cat gpurun_out/r2_pose_tail_prof.json; tail -5 gpurun_out/r2_pose_tail_prof.err
