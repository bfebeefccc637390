cd $GRAFT_REPO_ROOT
# eight GPUs: the bench line on the final code
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29515 bench.py --gpus 8 --steps 10 --warmup 3 --no-cpu-baseline --no-strict > gpurun_out/r2t_bench8.json 2> gpurun_out/r2t_bench8.err; echo bench8 rc=$?
tail -c 300 gpurun_out/r2t_bench8.err; grep '^{' gpurun_out/r2t_bench8.json | cut -c1-200
