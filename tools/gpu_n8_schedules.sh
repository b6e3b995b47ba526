#!/bin/bash
# H on 8 GPUs: the two exchange schedules back to back on the same box.
cd "$(dirname "$0")/.."
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29591"
for S in update 1; do
HPF_EXCHANGE_OVERLAP=$S HPF_MULTI=nvls HPF_GRAPH=1 timeout 120 $TR bench.py --gpus 8 --steps 20 --warmup 5 --no-cpu-baseline --no-e2e --no-parity-check > gpurun_out/bench_N8_sched_$S.json 2> gpurun_out/bench_N8_sched_$S.err
echo "schedule=$S $(grep '^{' gpurun_out/bench_N8_sched_$S.json | cut -c1-200)"
done
