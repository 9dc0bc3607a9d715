cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
SECONDS=0
CRDPN_BENCH_QUICK=1 timeout 200 python bench.py --steps 3 --warmup 3 > gpurun_out/quick_plain.json 2> gpurun_out/quick_plain.err; echo plain rc=$? ${SECONDS}s
cat gpurun_out/quick_plain.json
CRDPN_BENCH_QUICK=1 timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/r2_launches_bench_quick_final.csv python bench.py --steps 3 --warmup 3 > gpurun_out/ncu_quick.log 2>&1; echo ncu rc=$? ${SECONDS}s
grep -c gpu__time_duration gpurun_out/r2_launches_bench_quick_final.csv
