cd $GRAFT_REPO_ROOT
HPF_EXCHANGE_OVERLAP=update HPF_MULTI=nvls HPF_GRAPH=1 timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29581 bench.py --gpus 8 --config C5 --steps 10 --warmup 4 --no-cpu-baseline --no-e2e > gpurun_out/bench_C5_N8_update.json 2> gpurun_out/bench_C5_N8_update.err
grep '^{' gpurun_out/bench_C5_N8_update.json | cut -c1-230
tail -3 gpurun_out/bench_C5_N8_update.err
