#!/usr/bin/env python3
"""One-pass (fused gather + RED) sweep vs two-pass sweep, GPU box.
python tools/tune_fused.py [--alpha 0.6]"""
import argparse, json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch, bench
from hpfrec_b200.engine import Engine
from hpfrec_b200.loops import CudaLoops

ap = argparse.ArgumentParser()
ap.add_argument("--alpha", type=float, default=0.6)
ap.add_argument("--k", type=int, default=50)
a = ap.parse_args()
nU, nI, nnz, k = 1_000_000, 380_000, 48_000_000, a.k
dev = torch.device("cuda", 0)
u, i, y = bench.synth_coo_torch(nU, nI, nnz, dev, alpha=a.alpha)
print("max item degree", int(torch.bincount(i).max()), "max user degree", int(torch.bincount(u).max()))
u, i = u.to(torch.int32).contiguous(), i.to(torch.int32).contiguous()
loops = CudaLoops(True, device=0)
state = loops.initialize_parameters(np.empty((nU, k), np.float32), np.empty((nI, k), np.float32), 123, 0.3, 0.3, 1.0, 0.3, 0.3, 1.0)
eng = Engine(nU, nI, k, 4, 0)
eng.load_state(*state)
eng.set_option("panel_mb", 48)
eng.load_coo(u, i, y)

def run(tag, **opts):
    for n, v in opts.items():
        eng.set_option(n, v)
    eng.step_full(2)
    eng.set_option("timing", 1)
    eng.step_full(3)
    torch.cuda.synchronize()
    ms, n = eng.phase_ms()
    eng.set_option("timing", 0)
    print(json.dumps(dict(tag=tag, alpha=a.alpha, **opts, ms=[round(x / n, 3) for x in ms], total=round(sum(ms) / n, 3))), flush=True)

run("two-pass", sweep=0, lpg=4, unroll=1, minb=3, hint=1, chunk=64)
run("two-pass", sweep=0, lpg=8, unroll=1, minb=4, hint=0, chunk=128)
run("coo-atomic", sweep=1)
for lpg, minb, hint in ((4, 3, 0), (4, 3, 1), (8, 3, 0), (8, 3, 1), (8, 4, 0), (8, 4, 1), (4, 2, 0), (8, 2, 0)):
    for chunk in (64, 256):
        run("fused", sweep=2, lpg=lpg, minb=minb, hint=hint, chunk=chunk)
