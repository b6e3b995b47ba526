#!/bin/bash
# Validation of a changed default: full GPU test-suite, smoke(), then every single-GPU configuration with the default
# and with overlap_update=0.
cd "$(dirname "$0")/.."
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -3
timeout 200 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
for CFG in H C2 C3 H_f64; do
for O in -1 0; do
timeout 300 python bench.py --config $CFG --steps 20 --warmup 5 --no-cpu-baseline --option overlap_update=$O > gpurun_out/bench_${CFG}_ovu$O.json 2> gpurun_out/bench_${CFG}_ovu$O.err
echo "$CFG overlap_update=$O $(grep '^{' gpurun_out/bench_${CFG}_ovu$O.json | python -c 'import sys,json; d=json.loads(sys.stdin.read()); print(round(d["ms_per_step"],4), round(d["value"],1), "e2e", round(d["e2e"]["value"],1), "frac", round(d["roofline"]["frac"],4))')"
done
done
