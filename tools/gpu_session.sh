#!/bin/bash
# One GPU-box session (1 GPU): parity tests -> sweep tuner -> TMA staging probe -> ncu captures.
# Every step has its own timeout and writes under gpurun_out/, most valuable first, so a clamped call
# still brings results back.      gpurun --timeout 900 -- 'bash tools/gpu_session.sh [steps]'
set -u
STEPS=${1:-test,tune,probe,ncu}
cd "$(dirname "$0")/.."
OUT=gpurun_out
mkdir -p $OUT
T0=$(date +%s)
log() { echo "[+$(( $(date +%s) - T0 ))s] $*" | tee -a $OUT/session.log; }
has() { case ",$STEPS," in *",$1,"*) return 0;; *) return 1;; esac; }
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,clocks.mem,power.draw,clocks_event_reasons.active --format=csv > $OUT/smi_start.csv 2>&1
python -c "import os; print('cpus', len(os.sched_getaffinity(0)))" > $OUT/host.txt 2>&1

if has test; then
log "parity tests"
timeout 900 python -m pytest tests -m gpu -q ${PYTEST_ARGS:-} > $OUT/pytest_parity.log 2>&1
log "  rc=$? $(tail -1 $OUT/pytest_parity.log)"
fi
if has tune; then
log "tuner"
timeout 480 python tools/tune.py --budget-s ${TUNE_BUDGET:-240} ${TUNE_ARGS:-} > $OUT/tune.log 2>&1
log "  rc=$? $(tail -1 $OUT/tune.log)"
fi
if has probe; then
log "TMA staging probe"
timeout 150 tools/bin/tma_gather_probe --rows 1000000 > $OUT/tma_gather_probe_users.jsonl 2> $OUT/tma_gather_probe.err
timeout 100 tools/bin/tma_gather_probe --rows 380000 > $OUT/tma_gather_probe_items.jsonl 2>> $OUT/tma_gather_probe.err
log "  rc=$? $(wc -l < $OUT/tma_gather_probe_users.jsonl) + $(wc -l < $OUT/tma_gather_probe_items.jsonl) lines"
fi
if has ncu; then
for CFG in ${NCU_CFGS:-"chunk=256"}; do
TAG=$(echo $CFG | tr ',=' '__')
log "ncu --set full, $CFG"
OPTS=""; for o in $(echo $CFG | tr ',' ' '); do OPTS="$OPTS --option $o"; done
timeout 240 ncu --set full --clock-control none --import-source on -k regex:'sweep_rows|update_rows' -s 8 -c 4 \
    -f -o $OUT/ncu_full_$TAG python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu-baseline $OPTS > $OUT/ncu_full_$TAG.log 2>&1
log "  rc=$?"
# gpurun_out/ is capped at 64 MiB: keep the raw and source pages as CSV, drop the report itself
ncu -i $OUT/ncu_full_$TAG.ncu-rep --page raw --csv > $OUT/ncu_full_$TAG.raw.csv 2>/dev/null
ncu -i $OUT/ncu_full_$TAG.ncu-rep --page source --csv -k regex:sweep_rows -c 1 > $OUT/ncu_full_$TAG.source.csv 2>/dev/null
rm -f $OUT/ncu_full_$TAG.ncu-rep
done
fi
if has traffic; then
log "DRAM traffic probe (ncu metrics)"
timeout 400 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct,gpu__time_duration.sum \
    --clock-control none -k regex:'sweep_rows|update_rows' --csv --log-file $OUT/traffic.csv python tools/traffic_probe.py > $OUT/traffic_probe.log 2>&1
log "  rc=$? $(wc -l < $OUT/traffic.csv) csv lines"
fi
if has launches; then
log "ncu launch list"
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 500 --csv --log-file $OUT/launches.csv \
    python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu-baseline > $OUT/launches_bench.log 2>&1
log "  rc=$?"
fi
if has bench; then
log "bench"
timeout 400 python bench.py --steps 20 --warmup 3 ${BENCH_ARGS:-} > $OUT/bench.json 2> $OUT/bench.err
log "  rc=$? $(cut -c1-200 $OUT/bench.json)"
fi
if has c4; then
log "bench --config C4 (SVI epochs)"
timeout 300 python bench.py --config C4 --steps 8 --warmup 4 > $OUT/bench_C4.json 2> $OUT/bench_C4.err
log "  rc=$? $(cut -c1-200 $OUT/bench_C4.json)"
fi
if has c5smoke; then
log "bench --config C5 on ONE GPU (plumbing check of the 500M-nnz path)"
timeout 600 python bench.py --config C5 --steps 3 --warmup 3 --no-e2e --no-cpu-baseline > $OUT/bench_C5_N1.json 2> $OUT/bench_C5_N1.err
log "  rc=$? $(cut -c1-300 $OUT/bench_C5_N1.json)"
fi
if has c4launches; then
log "ncu launch list of two SVI epochs"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file $OUT/launches_C4.csv \
    python bench.py --config C4 --steps 2 --warmup 2 --no-e2e --no-cpu-baseline > $OUT/launches_C4.log 2>&1
log "  rc=$?"
fi
if has configs; then
for CFG in C2 C3 H_f64; do
log "bench --config $CFG"
timeout 300 python bench.py --config $CFG --steps 20 --warmup 3 --no-cpu-baseline > $OUT/bench_$CFG.json 2> $OUT/bench_$CFG.err
log "  rc=$? $(cut -c1-200 $OUT/bench_$CFG.json)"
done
fi
if has refarm; then
log "bench --impl reference (full configuration, ${REF_STEPS:-3} steps)"
timeout 900 python bench.py --impl reference --steps ${REF_STEPS:-3} --warmup ${REF_WARMUP:-1} > $OUT/bench_reference.json 2> $OUT/bench_reference.err
log "  rc=$? $(cut -c1-300 $OUT/bench_reference.json)"
fi
if has fullsize; then
log "full-size invariants"
timeout 400 python -m pytest tests/test_gpu_fullsize.py -m gpu -x -q > $OUT/pytest_fullsize.log 2>&1
log "  rc=$? $(tail -1 $OUT/pytest_fullsize.log)"
fi
log "done"
