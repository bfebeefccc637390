cd $GRAFT_REPO_ROOT
timeout 120 python dev/visc_param_ab.py cuda 64 mg_xgroup 0 1 > gpurun_out/r2n_xg_ab.log 2>&1; echo ab rc=$?
timeout 200 python tests/gpu_dev_gmg.py 256 2 4 mg_xgroup=1 > gpurun_out/r2n_xg1.log 2>&1; echo x1 rc=$?
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2n_launches.csv python tests/gpu_dev_gmg.py 256 2 2 > gpurun_out/r2n_ncu.log 2>&1; echo ncu rc=$?
python tests/gpu_dev_launchlist.py gpurun_out/r2n_launches.csv 70 > gpurun_out/r2n_launches.txt; rm -f gpurun_out/r2n_launches.csv
tail -n 5 gpurun_out/r2n_xg_ab.log gpurun_out/r2n_xg1.log; head -12 gpurun_out/r2n_launches.txt
