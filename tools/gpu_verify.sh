#!/bin/bash
# Shortest end-of-round check under the shipped defaults: what the driver runs (GPU tests, smoke, bench).
#     gpurun --timeout 300 -- 'bash tools/gpu_verify.sh'
set -u
cd "$(dirname "$0")/.."
OUT=gpurun_out
mkdir -p $OUT
T0=$(date +%s)
log() { echo "[+$(( $(date +%s) - T0 ))s] $*" | tee -a $OUT/verify.log; }
timeout 200 python -m pytest tests -m gpu -x -q > $OUT/pytest_verify.log 2>&1
log "pytest rc=$? $(tail -1 $OUT/pytest_verify.log)"
timeout 60 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/smoke_verify.log 2>&1
log "smoke rc=$? $(tail -1 $OUT/smoke_verify.log)"
timeout 120 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > $OUT/bench_verify.json 2> $OUT/bench_verify.err
log "bench rc=$? $(cut -c1-140 $OUT/bench_verify.json)"
