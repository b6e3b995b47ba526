#!/bin/bash
# overlap_update (user update under the item-major pass): parity test, then H for several side-launch shapes.
#   CONFIGS="ctas:block:pipe ..."
cd "$(dirname "$0")/.."
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "under_the_item_pass" 2>&1 | tail -2
for C in ${CONFIGS:-0:128:1 296:128:1 148:128:1 222:128:1 296:128:0 444:64:1 592:64:1 888:64:1 0:128:2}; do
IFS=: read N B P <<< "$C"
timeout 200 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-e2e --option overlap_update=$N --option overlap_block=$B --option update_pipe=$P ${BENCH_EXTRA:-} > gpurun_out/bench_ovu_$N-$B-$P.json 2> gpurun_out/bench_ovu_$N-$B-$P.err
echo "ctas=$N block=$B pipe=$P $(grep '^{' gpurun_out/bench_ovu_$N-$B-$P.json | python -c 'import sys,json; d=json.loads(sys.stdin.read()); print(round(d["ms_per_step"],4), [round(k["ms"],3) for k in d["roofline"]["kernels"]])')"
done
