cd $GRAFT_REPO_ROOT
timeout 120 python dev/visc_param_ab.py cuda 64 mg_tail 0 1 > gpurun_out/r2l_tail_ab.log 2>&1; echo ab rc=$?
timeout 200 python tests/gpu_dev_gmg.py 256 2 4 mg_tail=1 > gpurun_out/r2l_tail1.log 2>&1; echo t1 rc=$?
timeout 200 python tests/gpu_dev_gmg.py 256 2 4 mg_tail=0 pressure_resident=0 > gpurun_out/r2l_tail0.log 2>&1; echo t0 rc=$?
timeout 300 python bench.py --steps 10 --warmup 3 --no-strict --no-cpu-baseline > gpurun_out/r2l_bench.json 2> gpurun_out/r2l_bench.err; echo bench rc=$?
tail -n 5 gpurun_out/r2l_tail_ab.log gpurun_out/r2l_tail1.log gpurun_out/r2l_tail0.log; tail -c 600 gpurun_out/r2l_bench.err
