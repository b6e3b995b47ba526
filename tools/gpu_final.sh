#!/bin/bash
# Final single-GPU session of a round: (A) the vector-triple kernel against the shipped default, tests and a
# full bench under it; (B) everything the driver runs at round end, under the shipped defaults.
#     gpurun --timeout 600 -- 'bash tools/gpu_final.sh'
set -u
cd "$(dirname "$0")/.."
OUT=gpurun_out
mkdir -p $OUT
T0=$(date +%s)
log() { echo "[+$(( $(date +%s) - T0 ))s] $*" | tee -a $OUT/final.log; }
log "A1 tune (v3 default vs v4)"
timeout 300 python tools/tune_v2.py --skip-v2 > $OUT/tune_v4.log 2>&1
log "  rc=$?"
eval "$(python tools/best_env.py H_k50_alpha0.6_v4)"
log "  v4 best: HPF_ROW_ALIGN=${HPF_ROW_ALIGN:-} HPF_OPTIONS=${HPF_OPTIONS:-}"
log "A2 parity + API tests under the v4 best"
timeout 300 python -m pytest tests/test_gpu_parity.py tests/test_gpu_api.py -m gpu -x -q > $OUT/pytest_v4.log 2>&1
log "  rc=$? $(tail -1 $OUT/pytest_v4.log)"
log "A3 bench under the v4 best"
timeout 300 python bench.py --steps 20 --warmup 3 > $OUT/bench_v4.json 2> $OUT/bench_v4.err
log "  rc=$? $(cut -c1-150 $OUT/bench_v4.json)"
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 500 --csv --log-file $OUT/launches_v4.csv \
    python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu-baseline > /dev/null 2>&1
unset HPF_ROW_ALIGN HPF_OPTIONS

log "B1 all GPU tests, shipped defaults"
timeout 500 python -m pytest tests -m gpu -x -q > $OUT/pytest_final.log 2>&1
log "  rc=$? $(tail -1 $OUT/pytest_final.log)"
log "B2 smoke()"
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/smoke.log 2>&1
log "  rc=$? $(tail -1 $OUT/smoke.log)"
log "B3 bench, shipped defaults"
timeout 300 python bench.py --steps 20 --warmup 3 > $OUT/bench_final.json 2> $OUT/bench_final.err
log "  rc=$? $(cut -c1-150 $OUT/bench_final.json)"
log "B4 ncu launch list, shipped defaults"
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 500 --csv --log-file $OUT/launches_final.csv \
    python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu-baseline > /dev/null 2>&1
log "  rc=$?"
log "B5 reference arm"
timeout 200 python bench.py --impl reference --steps 3 --warmup 1 > $OUT/bench_reference.json 2> $OUT/bench_reference.err
log "  rc=$? $(cut -c1-120 $OUT/bench_reference.json)"
log "B6 fp64 (parity instantiation): pipelined vs classic kernel"
timeout 120 python bench.py --dtype f64 --steps 5 --warmup 3 --no-e2e --no-cpu-baseline > $OUT/bench_f64_k3.json 2>/dev/null
timeout 120 python bench.py --dtype f64 --steps 5 --warmup 3 --no-e2e --no-cpu-baseline --option kernel=1 --option chunk=64 > $OUT/bench_f64_k1.json 2>/dev/null
log "  $(cut -c60-130 $OUT/bench_f64_k3.json) | $(cut -c60-130 $OUT/bench_f64_k1.json)"
log "done"
