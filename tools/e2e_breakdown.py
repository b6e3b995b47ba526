#!/usr/bin/env python3
"""Where the end-to-end (host buffers) call spends its time: create / load_state / load_coo / iterate / export."""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch, bench
from hpfrec_b200.engine import Engine
from hpfrec_b200.loops import CudaLoops

nU, nI, nnz, k, K = 1_000_000, 380_000, 48_000_000, 50, 20
dev = torch.device("cuda", 0)
u, i, y = bench.synth_coo_torch(nU, nI, nnz, dev)
hu = u.to(torch.int32).cpu().pin_memory().numpy(); hi = i.to(torch.int32).cpu().pin_memory().numpy(); hy = y.cpu().pin_memory().numpy()
del u, i, y; torch.cuda.empty_cache()
loops = CudaLoops(True, device=0)
st = loops.initialize_parameters(np.empty((nU, k), np.float32), np.empty((nI, k), np.float32), 123, 0.3, 0.3, 1.0, 0.3, 0.3, 1.0)
hstate = [torch.from_numpy(x).pin_memory().numpy() for x in st]
outs = dict(Gamma_shp=(nU, k), Gamma_rte=(nU, k), Lambda_shp=(nI, k), Lambda_rte=(nI, k), k_rte=(nU, 1), t_rte=(nI, 1), Theta=(nU, k), Beta=(nI, k))
out = {key: torch.empty(shape, dtype=torch.float32).pin_memory().numpy() for key, shape in outs.items()}
for rep in range(3):
    t = [time.time()]
    def lap():
        torch.cuda.synchronize(); t.append(time.time())
    e = Engine(nU, nI, k, 4, 0); lap()
    e.load_state(*hstate); lap()
    e.load_coo(hu, hi, hy); lap()
    e.step_full(K); lap()
    e.export_state(**out); lap()
    e.close(); lap()
    names = ["create", "load_state", "load_coo", "iterate_%d" % K, "export", "destroy"]
    print(json.dumps({n: round(1e3 * (t[j + 1] - t[j]), 2) for j, n in enumerate(names)} | {"total_ms": round(1e3 * (t[-1] - t[0]), 2)}), flush=True)
