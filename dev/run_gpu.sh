cd $GRAFT_REPO_ROOT
timeout 200 python tests/gpu_dev_gmg.py 256 2 4 pressure_resident=0 > gpurun_out/r2i_pres0.log 2>&1; echo p0 rc=$?
timeout 200 python tests/gpu_dev_gmg.py 256 2 4 pressure_resident=1 > gpurun_out/r2i_pres1.log 2>&1; echo p1 rc=$?
timeout 300 python -m pytest tests/test_stage_parity.py tests/test_edge_cases.py tests/test_configs_gpu.py -x -q -m gpu > gpurun_out/r2i_tests.log 2>&1; echo tests rc=$?
tail -5 gpurun_out/r2i_pres0.log gpurun_out/r2i_pres1.log; tail -3 gpurun_out/r2i_tests.log
