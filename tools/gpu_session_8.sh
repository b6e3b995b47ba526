#!/bin/bash
# The 8-GPU session: exchange parity at 8 ranks, H at 8 and 4 GPUs (reduce-scatter overlap on/off, symmetric-memory
# vs NCCL barriers), then C5.      gpurun --gpus 8 --timeout 600 -- 'bash tools/gpu_session_8.sh'
set -u
cd "$(dirname "$0")/.."
OUT=gpurun_out
mkdir -p $OUT
T0=$(date +%s)
log() { echo "[+$(( $(date +%s) - T0 ))s] $*" | tee -a $OUT/session_8.log; }
run() {  # run <tag> <nproc> <env...> -- <bench args...>
    local tag=$1 n=$2; shift 2
    local envs=()
    while [ "$1" != "--" ]; do envs+=("$1"); shift; done
    shift
    log "bench $tag"
    env "${envs[@]}" timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29551 \
        bench.py --gpus $n "$@" > $OUT/bench_$tag.json 2> $OUT/bench_$tag.err
    log "  rc=$? $(grep '^{' $OUT/bench_$tag.json | tail -1 | cut -c1-200)"
}
nvidia-smi --query-gpu=index,name,clocks.sm,power.draw --format=csv > $OUT/smi_8.csv 2>&1
log "exchange parity, 8 ranks"
HPF_TEST_MODES=nvls timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29550 \
    tests/multi_rank_worker.py > $OUT/parity_N8.log 2>&1
log "  rc=$? $(grep -c '^PARITY' $OUT/parity_N8.log) lines"
grep '^PARITY' $OUT/parity_N8.log | cut -c1-220 | tee -a $OUT/session_8.log
run N8_default 8 HPF_MULTI=nvls HPF_GRAPH=1 -- --steps 20 --warmup 5 --no-cpu-baseline
run N8_o1_nccl 8 HPF_MULTI=nvls HPF_GRAPH=1 HPF_SYNC=nccl -- --steps 20 --warmup 5 --no-cpu-baseline --no-e2e --no-parity-check
run N8_o0_nccl 8 HPF_MULTI=nvls HPF_GRAPH=1 HPF_SYNC=nccl HPF_EXCHANGE_OVERLAP=0 -- --steps 20 --warmup 5 --no-cpu-baseline --no-e2e --no-parity-check
run C5_N8_default 8 HPF_MULTI=nvls HPF_GRAPH=1 -- --config C5 --steps 10 --warmup 3 --no-cpu-baseline --no-e2e
run N4_default 4 HPF_MULTI=nvls HPF_GRAPH=1 -- --steps 20 --warmup 5 --no-cpu-baseline --no-e2e --no-parity-check
log "done"
