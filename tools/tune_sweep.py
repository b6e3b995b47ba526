#!/usr/bin/env python3
"""Sweep-kernel tuning harness (GPU box).  

    python -m hpfrec_b200.build                     (build container)
    python tools/tune_sweep.py [--k 50]  (GPU box, via gpurun)

Times the two sweep passes (CUDA events around each kernel, engine option "timing") for every
compiled (lane-group width, unroll, min blocks/SM, L2 hint) variant, then varies the L2 panel size and
the chunk length for the best few.  Results: gpurun_out/tune_k<k>.jsonl (one JSON object per line).
"""
import argparse
import itertools
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import torch  # noqa: E402

import bench  # noqa: E402
from hpfrec_b200.engine import Engine  # noqa: E402
from hpfrec_b200.loops import CudaLoops  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--nusers", type=int, default=1_000_000)
    ap.add_argument("--nitems", type=int, default=380_000)
    ap.add_argument("--nnz", type=int, default=48_000_000)
    ap.add_argument("--k", type=int, default=50)
    ap.add_argument("--iters", type=int, default=3)
    ap.add_argument("--quick", action="store_true")
    a = ap.parse_args()
    dev = torch.device("cuda", 0)
    u, i, y = bench.synth_coo_torch(a.nusers, a.nitems, a.nnz, dev)
    u, i = u.to(torch.int32).contiguous(), i.to(torch.int32).contiguous()
    loops = CudaLoops(True, device=0)
    Theta = np.empty((a.nusers, a.k), np.float32)
    Beta = np.empty((a.nitems, a.k), np.float32)
    state = loops.initialize_parameters(Theta, Beta, 123, 0.3, 0.3, 1.0, 0.3, 0.3, 1.0)
    eng = Engine(a.nusers, a.nitems, a.k, 4, 0)
    eng.load_state(*state)
    packs = eng.ld // 4
    bucket = next((b for b in (8, 16, 32) if packs <= b), None)
    lpgs = {8: [4, 8], 16: [4, 8, 16], 32: [8, 16, 32]}.get(bucket)
    if lpgs is None:
        raise SystemExit("no tuning variants compiled for ld=%d" % eng.ld)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    out = open(os.path.join(ROOT, "gpurun_out", "tune_k%d.jsonl" % a.k), "w")

    def measure(panel_mb, chunk, lpg, unroll, minb, hint, reload=False):
        if reload:
            eng.set_option("panel_mb", panel_mb)
            eng.load_coo(u, i, y)
        for name, val in (("chunk", chunk), ("lpg", lpg), ("unroll", unroll), ("minb", minb), ("hint", hint)):
            eng.set_option(name, val)
        eng.step_full(1)
        eng.set_option("timing", 1)
        eng.step_full(a.iters)
        torch.cuda.synchronize()
        ms, n = eng.phase_ms()
        eng.set_option("timing", 0)
        rec = dict(k=a.k, panel_mb=panel_mb, chunk=chunk, lpg=lpg, unroll=unroll, minb=minb, hint=hint,
                   ms_item_major=ms[0] / n, ms_user_major=ms[1] / n, ms_upd_users=ms[2] / n, ms_upd_items=ms[3] / n)
        rec["ms_sweep"] = rec["ms_item_major"] + rec["ms_user_major"]
        out.write(json.dumps(rec) + "\n")
        out.flush()
        return rec

    results = []
    first = True
    grid = list(itertools.product(lpgs, [1], [2, 3, 4], [0, 1, 3]))  # unrolled shapes were measured slower and removed
    if a.quick:
        grid = [g for g in grid if g[2] in (2, 4)]
    for lpg, unroll, minb, hint in grid:
        results.append(measure(48.0, 128, lpg, unroll, minb, hint, reload=first))
        first = False
    results.sort(key=lambda r: r["ms_sweep"])
    print("== top variants (panel 48 MB, chunk 128) ==")
    for r in results[:8]:
        print(json.dumps(r))
    top = results[:3]
    print("== panel / chunk sweep ==")
    best = []
    for panel in (12.0, 24.0, 32.0, 48.0, 64.0, 96.0, 100000.0):
        reload = True
        for r0 in top:
            for chunk in (64, 128, 256, 512):
                rec = measure(panel, chunk, r0["lpg"], r0["unroll"], r0["minb"], r0["hint"], reload=reload)
                reload = False
                best.append(rec)
    best.sort(key=lambda r: r["ms_sweep"])
    for r in best[:10]:
        print(json.dumps(r))
    out.close()


if __name__ == "__main__":
    main()
