set -x
cd $GRAFT_REPO_ROOT
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -30 > gpurun_out/r2_pytest5.log
tail -12 gpurun_out/r2_pytest5.log
python profiles/r2_shard_ab.py > gpurun_out/r2_shard_ab3.json 2>gpurun_out/r2_shard_ab3.err; python -c "
import json; d=json.load(open('gpurun_out/r2_shard_ab3.json'))
for k,v in d.items(): print(k, {a: round(b,4) for a,b in v.items()})"
tail -3 gpurun_out/r2_shard_ab3.err
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r2_bench5.json 2> gpurun_out/r2_bench5.err; echo bench rc=$?
tail -c 400 gpurun_out/r2_bench5.err
