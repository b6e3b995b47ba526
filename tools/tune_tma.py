#!/usr/bin/env python3
"""Staged-gather (cp.async.bulk) sweep vs register-gather sweep: agreement + timing."""
import argparse, json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch, bench
from hpfrec_b200.engine import Engine
from hpfrec_b200.loops import CudaLoops

ap = argparse.ArgumentParser()
ap.add_argument("--k", type=int, default=50)
ap.add_argument("--alpha", type=float, default=0.6)
a = ap.parse_args()
nU, nI, nnz, k = 1_000_000, 380_000, 48_000_000, a.k
dev = torch.device("cuda", 0)
u, i, y = bench.synth_coo_torch(nU, nI, nnz, dev, alpha=a.alpha)
u, i = u.to(torch.int32).contiguous(), i.to(torch.int32).contiguous()
loops = CudaLoops(True, device=0)
state = loops.initialize_parameters(np.empty((nU, k), np.float32), np.empty((nI, k), np.float32), 123, 0.3, 0.3, 1.0, 0.3, 0.3, 1.0)

def fresh(**opts):
    e = Engine(nU, nI, k, 4, 0)
    e.load_state(*state)
    e.set_option("panel_mb", 48)
    for n, v in opts.items():
        e.set_option(n, v)
    e.load_coo(u, i, y)
    return e

# agreement after 2 iterations
e0 = fresh(sweep=0); e0.step_full(2); r0 = e0.export_all(); e0.close()
e3 = fresh(sweep=3); e3.step_full(2); r3 = e3.export_all(); e3.close()
for key in ("Theta", "Beta", "k_rte", "t_rte"):
    err = float(np.max(np.abs(r0[key] - r3[key]) / np.maximum(np.abs(r0[key]), 1e-30)))
    print("agreement", key, "max rel diff %.3e" % err, flush=True)
    assert err < 1e-4, key

eng = fresh()
def run(tag, **opts):
    for n, v in opts.items():
        eng.set_option(n, v)
    eng.step_full(2)
    eng.set_option("timing", 1)
    eng.step_full(3)
    torch.cuda.synchronize()
    ms, n = eng.phase_ms()
    eng.set_option("timing", 0)
    print(json.dumps(dict(tag=tag, k=k, **opts, ms=[round(x / n, 3) for x in ms], sweep_ms=round((ms[0] + ms[1]) / n, 3))), flush=True)

run("reg-gather", sweep=0, lpg=4, unroll=1, minb=3, hint=1, chunk=64)
for minb in (2, 3, 4):
    run("reg-gather hint2", sweep=0, lpg=4, unroll=1, minb=minb, hint=2, chunk=64)
    run("reg-gather hint2", sweep=0, lpg=8, unroll=1, minb=minb, hint=2, chunk=64)
for minb in (2, 3, 4):
    for chunk in (64, 128, 256, 512, 1024):
        run("tma-staged", sweep=3, minb=minb, chunk=chunk)
