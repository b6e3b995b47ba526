#!/usr/bin/env python3
"""Host-overhead probe for the sharded loop (torchrun, small per-GPU problem): eager vs CUDA-graph replay."""
import os, sys, time
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
nU, nI, nnz, k = 125_000 * world, 380_000, 6_000_000 * world, 50
u, i, y = bench.synth_coo_torch(nU, nI, nnz, dev, seed=3)
loops = CudaLoops(True, device=local)
Gs, Gr, Ls, Lr, kr, tr = loops.initialize_parameters(np.empty((nU, k), np.float32), np.empty((nI, k), np.float32), 123, 0.3, 0.3, 1.0, 0.3, 0.3, 1.0)
cuts = hdist.plan_user_shards(u, nU, world); lo, hi = cuts[rank], cuts[rank + 1]
lu, li, ly = (t.contiguous() for t in hdist.shard_triples(u, i, y, lo, hi))

def make():
    e = Engine(hi - lo, nI, k, 4, local)
    e.load_state(np.ascontiguousarray(Gs[lo:hi]), np.ascontiguousarray(Gr[lo:hi]), Ls, Lr, np.ascontiguousarray(kr[lo:hi]), tr)
    e.load_coo(lu, li, ly)
    return e

def timeit(fn, n=30):
    fn(3); torch.cuda.synchronize(); dist.barrier()
    t0 = time.time(); fn(n); t_enq = time.time() - t0
    torch.cuda.synchronize(); dist.barrier(); t = time.time() - t0
    return 1e3 * t / n, 1e3 * t_enq / n

for mode in ("peer", "overlap"):
    e = make()
    if mode == "peer":
        hdist.attach_peers(e)
        ms, enq = timeit(lambda n: hdist.run_sharded_iterations_peer(e, n))
    else:
        ms, enq = timeit(lambda n: hdist.run_sharded_iterations_overlapped(e, n))
    if rank == 0:
        print("eager  %-8s %.3f ms/iter (host enqueue %.3f ms/iter)" % (mode, ms, enq), flush=True)
    ref = e.export_all()["Beta"]
    e.close()
    if os.environ.get("SKIP_GRAPH"):
        continue
    e = make()
    loop = hdist.GraphedShardLoop(e, mode)
    if rank == 0: print("capturing", mode, flush=True)
    with torch.cuda.stream(loop.stream):
        ms, enq = timeit(lambda n: loop.run(n))
    if rank == 0:
        print("graph  %-8s %.3f ms/iter (host enqueue %.3f ms/iter)" % (mode, ms, enq), flush=True)
    e.close()
dist.destroy_process_group()
