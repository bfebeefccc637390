cd $GRAFT_REPO_ROOT
# last call of the round: -m gpu suite on the final code, then timing of the reworked level-1 sweep
timeout 300 python -m pytest tests -x -q -m gpu > gpurun_out/r2v_tests.log 2>&1; echo tests rc=$?
timeout 60 python tests/gpu_dev_gmg.py 256 2 3 > gpurun_out/r2v_sweep.log 2>&1; echo sweep rc=$?
timeout 100 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-strict > gpurun_out/r2v_bench.json 2> gpurun_out/r2v_bench.err; echo bench rc=$?
tail -n 3 gpurun_out/r2v_tests.log; tail -n 3 gpurun_out/r2v_sweep.log; cut -c1-200 gpurun_out/r2v_bench.json
