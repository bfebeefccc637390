cd $GRAFT_REPO_ROOT
# four GPUs: the bench line on the final code
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 4 --steps 10 --warmup 3 --no-cpu-baseline --no-strict > gpurun_out/r2s_bench4.json 2> gpurun_out/r2s_bench4.err; echo bench4 rc=$?
tail -c 300 gpurun_out/r2s_bench4.err; grep '^{' gpurun_out/r2s_bench4.json | cut -c1-200
