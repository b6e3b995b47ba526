#!/usr/bin/env python3
"""Per-phase CUDA-event timing of the sharded iteration on every rank (torchrun; peer mode, eager)."""
import os, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch, torch.distributed as dist
from hpfrec_b200 import dist as hdist
from hpfrec_b200.engine import Engine
from hpfrec_b200.loops import CudaLoops
import bench

world = int(os.environ.get("WORLD_SIZE", "1")); rank = int(os.environ.get("RANK", "0")); local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local); dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
nU, nI, nnz, k = 1_000_000, 380_000, 48_000_000, 50
u, i, y = bench.synth_coo_torch(nU, nI, nnz, dev)
loops = CudaLoops(True, device=local)
Gs, Gr, Ls, Lr, kr, tr = loops.initialize_parameters(np.empty((nU, k), np.float32), np.empty((nI, k), np.float32), 123, 0.3, 0.3, 1.0, 0.3, 0.3, 1.0)
cuts = hdist.plan_user_shards(u, nU, world); lo, hi = cuts[rank], cuts[rank + 1]
lu, li, ly = (t.contiguous() for t in hdist.shard_triples(u, i, y, lo, hi))
del u, i, y
e = Engine(hi - lo, nI, k, 4, local)
e.load_state(np.ascontiguousarray(Gs[lo:hi]), np.ascontiguousarray(Gr[lo:hi]), Ls, Lr, np.ascontiguousarray(kr[lo:hi]), tr)
e.load_coo(lu, li, ly)
hdist.attach_peers(e)
_, _, p_theta, n_theta = e.partials(); p_beta, n_beta = e.beta_colsum()
t_theta = hdist.wrap_device_buffer(p_theta, n_theta, torch.float64, local)
t_beta = hdist.wrap_device_buffer(p_beta, n_beta, torch.float64, local)
names = ["pass_items", "pass_users", "update_users", "allreduce_theta", "update_items_peer", "allreduce_beta", "finish_memset"]
acc = np.zeros(len(names)); N = 20
s = torch.cuda.current_stream()
for it in range(N + 3):
    evs = [torch.cuda.Event(enable_timing=True) for _ in range(len(names) + 1)]
    evs[0].record(s); e.sweep_side(0)
    evs[1].record(s); e.sweep_side(1)
    evs[2].record(s); e.update_users()
    evs[3].record(s); dist.all_reduce(t_theta)
    evs[4].record(s); e.update_items_peer(False)
    evs[5].record(s); dist.all_reduce(t_beta)
    evs[6].record(s); e.peer_finish()
    evs[7].record(s)
    torch.cuda.synchronize()
    if it >= 3:
        acc += np.array([evs[j].elapsed_time(evs[j + 1]) for j in range(len(names))])
acc /= N
allp = [None] * world
dist.all_gather_object(allp, (rank, hi - lo, int(ly.shape[0]), acc.tolist()))
if rank == 0:
    for r, nu, nz, a in allp:
        print("rank %d users=%d nnz=%d " % (r, nu, nz) + " ".join("%s=%.3f" % (n, v) for n, v in zip(names, a)) + " total=%.3f" % sum(a), flush=True)
dist.destroy_process_group()
