#!/bin/bash
# Phase breakdown of the sharded loop on 8 GPUs (eager loop, HPF_PHASES=1), H and C5.
#     gpurun --gpus 8 --timeout 300 -- 'bash tools/gpu_phases_n8.sh'
cd "$(dirname "$0")/.."
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29571"
HPF_PHASES=1 HPF_GRAPH=0 HPF_MULTI=nvls timeout 120 $TR bench.py --gpus 8 --steps 20 --warmup 5 --no-cpu-baseline --no-e2e --no-parity-check > gpurun_out/phases_N8_H.log 2>&1
grep PHASES gpurun_out/phases_N8_H.log | tail -8 | cut -c1-600
grep '^{' gpurun_out/phases_N8_H.log | cut -c1-200
if [ "${WITH_C5:-0}" = "1" ]; then
HPF_PHASES=1 HPF_GRAPH=0 HPF_MULTI=nvls timeout 150 $TR bench.py --gpus 8 --config C5 --steps 6 --warmup 3 --no-cpu-baseline --no-e2e --no-parity-check > gpurun_out/phases_N8_C5.log 2>&1
grep PHASES gpurun_out/phases_N8_C5.log | tail -8 | cut -c1-600
fi
