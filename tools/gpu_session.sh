#!/bin/bash
# One GPU-box session: tune -> bench / tests / profiles under the fastest verified configuration ->
# the same under the shipped defaults.  Every step has its own timeout and writes under gpurun_out/,
# most valuable first, so a clamped call still brings results back.
#     gpurun --timeout 840 -- 'bash tools/gpu_session.sh [tuner.py] [best.json key]'
set -u
TUNER=${1:-tune_r2.py}
KEY=${2:-H_k50_alpha0.6}
cd "$(dirname "$0")/.."
OUT=gpurun_out
mkdir -p $OUT
T0=$(date +%s)
log() { echo "[+$(( $(date +%s) - T0 ))s] $*" | tee -a $OUT/session.log; }
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,clocks.mem,power.draw,clocks_event_reasons.active --format=csv > $OUT/smi_start.csv 2>&1
python - <<'PY' > $OUT/host.txt 2>&1
import os
print("cpus", len(os.sched_getaffinity(0)))
PY

log "1 $TUNER"
TUNER_NAME=${TUNER%% *}
timeout 420 python tools/$TUNER > $OUT/${TUNER_NAME%.py}.log 2>&1
log "  rc=$?"
eval "$(python tools/best_env.py $KEY)"
log "  winner: HPF_ROW_ALIGN=${HPF_ROW_ALIGN:-} HPF_OPTIONS=${HPF_OPTIONS:-}"

log "2 bench under the winner"
timeout 300 python bench.py --steps 20 --warmup 3 > $OUT/bench_winner.json 2> $OUT/bench_winner.err
log "  rc=$? $(cut -c1-160 $OUT/bench_winner.json)"

log "3 parity + API tests under the winner"
timeout 400 python -m pytest tests/test_gpu_parity.py tests/test_gpu_api.py -m gpu -x -q > $OUT/pytest_winner.log 2>&1
log "  rc=$? $(tail -1 $OUT/pytest_winner.log)"

if [ ! -s profiles/r01b_gather_probe_items.jsonl ]; then
log "4 gather / RED ceiling probe"
timeout 120 tools/bin/gather_probe --rows 380000 > $OUT/gather_probe_items.jsonl 2> $OUT/gather_probe.err
timeout 90 tools/bin/gather_probe --rows 1000000 --quick > $OUT/gather_probe_users.jsonl 2>> $OUT/gather_probe.err
log "  rc=$? $(wc -l < $OUT/gather_probe_items.jsonl) + $(wc -l < $OUT/gather_probe_users.jsonl) lines"
fi

log "5 ncu launch list under the winner"
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 500 --csv --log-file $OUT/launches_winner.csv \
    python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu-baseline > $OUT/launches_bench.log 2>&1
log "  rc=$?"

log "6 ncu --set full of the sweep and update kernels under the winner"
timeout 240 ncu --set full --clock-control none --import-source on -k regex:'sweep_major|update_rows' -s 8 -c 4 \
    -f -o $OUT/ncu_full_winner python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu-baseline > $OUT/ncu_full.log 2>&1
log "  rc=$?"

unset HPF_ROW_ALIGN HPF_OPTIONS
if [ -n "${SKIP_SHIPPED:-}" ]; then log "done (shipped-defaults steps skipped)"; exit 0; fi
log "7 bench, shipped defaults"
timeout 240 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > $OUT/bench_shipped.json 2> $OUT/bench_shipped.err
log "  rc=$? $(cut -c1-160 $OUT/bench_shipped.json)"

log "8 parity + API tests, shipped defaults"
timeout 400 python -m pytest tests/test_gpu_parity.py tests/test_gpu_api.py -m gpu -x -q > $OUT/pytest_shipped.log 2>&1
log "  rc=$? $(tail -1 $OUT/pytest_shipped.log)"

log "9 full-size invariants, shipped defaults"
timeout 400 python -m pytest tests/test_gpu_fullsize.py -m gpu -x -q > $OUT/pytest_fullsize.log 2>&1
log "  rc=$? $(tail -1 $OUT/pytest_fullsize.log)"
log "done"
