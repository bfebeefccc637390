cd $GRAFT_REPO_ROOT
# two GPUs: sharded-vs-single tests and the bench line on the final code
timeout 500 python -m pytest tests/test_multigpu.py -x -q -m gpu > gpurun_out/r2r_mgpu_tests.log 2>&1; echo tests rc=$?
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2r_bench2.json 2> gpurun_out/r2r_bench2.err; echo bench2 rc=$?
tail -n 3 gpurun_out/r2r_mgpu_tests.log; tail -c 300 gpurun_out/r2r_bench2.err; grep '^{' gpurun_out/r2r_bench2.json | cut -c1-200
