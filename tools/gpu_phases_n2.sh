#!/bin/bash
# Phase breakdown of the sharded loop on 2 GPUs (eager, HPF_PHASES=1) for several reduce-scatter grid sizes, then the
# graph-replay bench.      gpurun --gpus 2 --timeout 400 -- 'bash tools/gpu_phases_n2.sh'
cd "$(dirname "$0")/.."
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29561"
for CT in ${RS_CTAS:-24 48 96}; do
HPF_OPTIONS="rs_ctas=$CT" HPF_PHASES=1 HPF_GRAPH=0 HPF_MULTI=nvls timeout 200 $TR bench.py --gpus 2 --steps 20 --warmup 5 --no-cpu-baseline --no-e2e --no-parity-check > gpurun_out/phases_N2_rs$CT.log 2>&1
echo "rs_ctas=$CT"; grep PHASES gpurun_out/phases_N2_rs$CT.log | tail -2
done
HPF_GRAPH=1 HPF_MULTI=nvls timeout 200 $TR bench.py --gpus 2 --steps 20 --warmup 5 --no-cpu-baseline --no-e2e > gpurun_out/bench_N2_default.json 2> gpurun_out/bench_N2_default.err
grep '^{' gpurun_out/bench_N2_default.json | cut -c1-220
