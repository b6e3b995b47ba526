#!/usr/bin/env python3
"""Config C4: SVI epochs at 1M x 380K x 48M nnz, k=50, users_per_batch=50k,
items_per_batch=20k on one B200 -- seconds per epoch with device-side minibatch assembly."""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch, bench
from hpfrec_b200.engine import Engine
from hpfrec_b200.loops import CudaLoops

nU, nI, nnz, k = 1_000_000, 380_000, 48_000_000, 50
upb, ipb = 50_000, 20_000
dev = torch.device("cuda", 0)
u, i, y = bench.synth_coo_torch(nU, nI, nnz, dev)
u, i = u.to(torch.int32).contiguous(), i.to(torch.int32).contiguous()
loops = CudaLoops(True, device=0)
state = loops.initialize_parameters(np.empty((nU, k), np.float32), np.empty((nI, k), np.float32), 123, 0.3, 0.3, 1.0, 0.3, 0.3, 1.0)
eng = Engine(nU, nI, k, 4, 0)
eng.load_state(*state)
eng.set_option("panel_mb", 1e9)
t0 = time.time(); eng.load_coo(u, i, y); torch.cuda.synchronize(); t_load = time.time() - t0
rng = np.random.default_rng(123)
un, inn = np.arange(nU, dtype=np.int64), np.arange(nI, dtype=np.int64)
res = []
for e in range(4):
    rho = float(np.float32(1 / np.sqrt(e + 2)))
    user_epoch = ((e + 1) % 2) == 0
    l0 = eng.launch_count
    torch.cuda.synchronize(); t0 = time.time()
    if user_epoch:
        rng.shuffle(un)
        nb = int(np.ceil(nU / upb))
        for bt in range(nb):
            ids = np.ascontiguousarray(un[bt * upb: min(nU, (bt + 1) * upb)])
            eng.step_batch_ids(ids, True, rho, nU / ids.shape[0], False)
    else:
        rng.shuffle(inn)
        nb = int(np.ceil(nI / ipb))
        for bt in range(nb):
            ids = np.ascontiguousarray(inn[bt * ipb: min(nI, (bt + 1) * ipb)])
            eng.step_batch_ids(ids, False, rho, nI / ids.shape[0], False)
    torch.cuda.synchronize(); dt = time.time() - t0
    res.append(dict(epoch=e, kind="users" if user_epoch else "items", minibatches=nb, seconds=round(dt, 4),
                    nnz_per_s=round(nnz / dt / 1e9, 3), launches=eng.launch_count - l0))
    print(json.dumps(res[-1]), flush=True)
out = eng.export_all()
print(json.dumps(dict(load_coo_s=round(t_load, 3), finite=bool(np.isfinite(out["Theta"]).all() and np.isfinite(out["Beta"]).all()))))
