#!/bin/bash
# Multi-GPU session: parity of every exchange mode (2 ranks and N ranks), then the bench at N GPUs.
#     gpurun --gpus N --timeout 900 -- 'bash tools/gpu_session_multi.sh N [steps]'
set -u
N=${1:-2}
STEPS=${2:-parity,bench}
cd "$(dirname "$0")/.."
OUT=gpurun_out
mkdir -p $OUT
T0=$(date +%s)
log() { echo "[+$(( $(date +%s) - T0 ))s] $*" | tee -a $OUT/session_multi.log; }
has() { case ",$STEPS," in *",$1,"*) return 0;; *) return 1;; esac; }
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
nvidia-smi --query-gpu=index,name,clocks.sm,power.draw --format=csv > $OUT/smi_multi.csv 2>&1
if has parity; then
log "exchange parity, $N ranks (modes: ${HPF_TEST_MODES:-default})"
timeout 300 $TR --master-port 29541 tests/multi_rank_worker.py > $OUT/parity_N$N.log 2>&1
log "  rc=$? $(grep -c '^PARITY' $OUT/parity_N$N.log) lines"
grep '^PARITY' $OUT/parity_N$N.log | cut -c1-330 | tee -a $OUT/session_multi.log
fi
if has bench; then
for MODE in ${BENCH_MODES:-nvls peer}; do
for GRAPH in ${BENCH_GRAPH:-1 0}; do
for OVL in ${BENCH_OVERLAP:-1 0}; do
for SYNC in ${BENCH_SYNC:-symm nccl}; do
log "bench N=$N exchange=$MODE graph=$GRAPH overlap=$OVL sync=$SYNC"
TAG=N${N}_${MODE}_g${GRAPH}_o${OVL}_$SYNC
HPF_SYNC=$SYNC HPF_MULTI=$MODE HPF_GRAPH=$GRAPH HPF_EXCHANGE_OVERLAP=$OVL timeout 300 $TR --master-port 29542 bench.py --gpus $N --steps 20 --warmup 5 --no-cpu-baseline ${BENCH_ARGS:-} \
    > $OUT/bench_$TAG.json 2> $OUT/bench_$TAG.err
log "  rc=$? $(grep '^{' $OUT/bench_$TAG.json | cut -c1-230)"
done
done
done
done
fi
if has c5; then
log "bench --config C5, N=$N"
HPF_MULTI=${C5_MODE:-nvls} HPF_GRAPH=${C5_GRAPH:-1} timeout 600 $TR --master-port 29543 bench.py --gpus $N --config C5 --steps 10 --warmup 3 --no-cpu-baseline --no-e2e \
    > $OUT/bench_C5_N$N.json 2> $OUT/bench_C5_N$N.err
log "  rc=$? $(cut -c1-300 $OUT/bench_C5_N$N.json)"
fi
log "done"
