set -x
cd $GRAFT_REPO_ROOT
python dev/gmg_build_ab.py cuda 64 > gpurun_out/r2h_build_ab.log 2>&1; echo ab rc=$?
python bench.py --steps 10 --warmup 3 --no-strict --no-cpu-baseline --param mg_build=0 > gpurun_out/r2h_bench_build0.json 2> gpurun_out/r2h_bench_build0.err; echo b0 rc=$?
python bench.py --steps 10 --warmup 3 --no-strict --no-cpu-baseline > gpurun_out/r2h_bench_build1.json 2> gpurun_out/r2h_bench_build1.err; echo b1 rc=$?
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2h_launches.csv python tests/gpu_dev_gmg.py 256 2 2 > gpurun_out/r2h_ncu.log 2>&1; echo ncu rc=$?
python tests/gpu_dev_launchlist.py gpurun_out/r2h_launches.csv 70 > gpurun_out/r2h_launches.txt; rm -f gpurun_out/r2h_launches.csv
tail -3 gpurun_out/r2h_build_ab.log
