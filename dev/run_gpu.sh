cd $GRAFT_REPO_ROOT
# final one-GPU validation: compact level-1 rows A/B, bench line, -m gpu suite, smoke
timeout 100 python tests/gpu_dev_gmg.py 256 2 3 mg_compact=0 > gpurun_out/r2u_compact0.log 2>&1; echo c0 rc=$?
timeout 100 python tests/gpu_dev_gmg.py 256 2 3 mg_compact=1 > gpurun_out/r2u_compact1.log 2>&1; echo c1 rc=$?
timeout 150 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2u_bench.json 2> gpurun_out/r2u_bench.err; echo bench rc=$?
timeout 60 python dev/visc_param_ab.py cuda 64 mg_compact 0 1 > gpurun_out/r2u_compact_ab.log 2>&1; echo ab rc=$?
timeout 400 python -m pytest tests -x -q -m gpu > gpurun_out/r2u_tests.log 2>&1; echo tests rc=$?
timeout 60 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2u_smoke.log 2>&1; echo smoke rc=$?
tail -n 3 gpurun_out/r2u_compact0.log gpurun_out/r2u_compact1.log gpurun_out/r2u_compact_ab.log; tail -n 3 gpurun_out/r2u_tests.log; tail -n 2 gpurun_out/r2u_smoke.log; tail -c 300 gpurun_out/r2u_bench.err; cut -c1-200 gpurun_out/r2u_bench.json
