cd $GRAFT_REPO_ROOT
# programmatic dependent launch of the CG + V-cycle chunk: A/B (one GPU)
timeout 120 python dev/visc_param_ab.py cuda 64 pdl 0 1 > gpurun_out/r2q_pdl_ab.log 2>&1; echo ab rc=$?
timeout 200 python tests/gpu_dev_gmg.py 256 2 5 pdl=0 > gpurun_out/r2q_pdl0.log 2>&1; echo p0 rc=$?
timeout 200 python tests/gpu_dev_gmg.py 256 2 5 pdl=1 > gpurun_out/r2q_pdl1.log 2>&1; echo p1 rc=$?
timeout 300 python bench.py --steps 10 --warmup 3 --param pdl=1 --no-cpu-baseline > gpurun_out/r2q_bench_pdl1.json 2> gpurun_out/r2q_bench_pdl1.err; echo bench rc=$?
tail -n 3 gpurun_out/r2q_pdl_ab.log; tail -n 4 gpurun_out/r2q_pdl0.log gpurun_out/r2q_pdl1.log; tail -c 400 gpurun_out/r2q_bench_pdl1.err; cut -c1-220 gpurun_out/r2q_bench_pdl1.json
