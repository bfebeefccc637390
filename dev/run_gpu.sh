cd $GRAFT_REPO_ROOT
timeout 120 python dev/visc_param_ab.py cuda 64 mg_xgroup 0 1 > gpurun_out/r2m_xg_ab.log 2>&1; echo ab rc=$?
timeout 200 python tests/gpu_dev_gmg.py 256 2 4 mg_xgroup=0 > gpurun_out/r2m_xg0.log 2>&1; echo x0 rc=$?
timeout 200 python tests/gpu_dev_gmg.py 256 2 4 mg_xgroup=1 > gpurun_out/r2m_xg1.log 2>&1; echo x1 rc=$?
timeout 300 python bench.py --steps 10 --warmup 3 --no-strict --no-cpu-baseline > gpurun_out/r2m_bench.json 2> gpurun_out/r2m_bench.err; echo bench rc=$?
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2m_launches.csv python tests/gpu_dev_gmg.py 256 2 2 > gpurun_out/r2m_ncu.log 2>&1; echo ncu rc=$?
python tests/gpu_dev_launchlist.py gpurun_out/r2m_launches.csv 70 > gpurun_out/r2m_launches.txt; rm -f gpurun_out/r2m_launches.csv
tail -n 5 gpurun_out/r2m_xg_ab.log gpurun_out/r2m_xg0.log gpurun_out/r2m_xg1.log; tail -c 400 gpurun_out/r2m_bench.err
