#!/usr/bin/env python3
"""Multi-GPU parity check (run under torchrun, one rank per GPU):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
        --master-port 29512 tools/check_multi_gpu.py

Every rank fits its user shard with the overlapped NCCL loop; rank 0 also runs the whole problem on
one engine and checks that the gathered sharded result equals it (fp64: <= 1e-10; the
item-side replicas must be bit-identical across ranks)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from hpfrec_b200 import dist as hdist  # noqa: E402
from hpfrec_b200.engine import Engine  # noqa: E402
from hpfrec_b200.loops import CudaLoops  # noqa: E402
import bench  # noqa: E402


def main():
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    ok = True
    for dtype, rb, tol, overlapped in ((np.float64, 8, 1e-10, True), (np.float64, 8, 1e-10, False),
                                       (np.float64, 8, 1e-10, "peer"), (np.float32, 4, 2e-4, True),
                                       (np.float32, 4, 2e-4, "peer")):
        nU, nI, nnz, k, its = 60000, 25000, 1_500_000, 50, 6
        u, i, y = bench.synth_coo_torch(nU, nI, nnz, dev, seed=7)
        y = y.to(torch.float64 if rb == 8 else torch.float32)
        loops = CudaLoops(rb == 4, device=local)
        Theta, Beta = np.empty((nU, k), dtype), np.empty((nI, k), dtype)
        Gs, Gr, Ls, Lr, kr, tr = loops.initialize_parameters(Theta, Beta, 123, 0.3, 0.3, 1.0, 0.3, 0.3, 1.0)
        cuts = hdist.plan_user_shards(u, nU, world)
        lo, hi = cuts[rank], cuts[rank + 1]
        lu, li, ly = (t.contiguous() for t in hdist.shard_triples(u, i, y, lo, hi))
        eng = Engine(hi - lo, nI, k, rb, local)
        eng.load_state(np.ascontiguousarray(Gs[lo:hi]), np.ascontiguousarray(Gr[lo:hi]), Ls, Lr,
                       np.ascontiguousarray(kr[lo:hi]), tr)
        eng.load_coo(lu, li, ly)
        if overlapped == "peer":
            hdist.attach_peers(eng)
            hdist.run_sharded_iterations_peer(eng, its)
        elif overlapped:
            hdist.run_sharded_iterations_overlapped(eng, its)
        else:
            hdist.run_sharded_iterations(eng, its)
        torch.cuda.synchronize()
        mine = eng.export_all()
        eng.close()
        # replicas identical?
        beta = torch.from_numpy(mine["Beta"]).to(dev)
        ref_beta = beta.clone()
        dist.broadcast(ref_beta, 0)
        same = bool(torch.equal(beta, ref_beta))
        # gather user side on rank 0
        gathered = [None] * world
        dist.all_gather_object(gathered, (lo, hi, mine["Theta"], mine["k_rte"]))
        if rank == 0:
            e1 = Engine(nU, nI, k, rb, local)
            e1.load_state(Gs, Gr, Ls, Lr, kr, tr)
            e1.load_coo(u.contiguous(), i.contiguous(), y.contiguous())
            e1.step_full(its)
            single = e1.export_all()
            e1.close()
            theta = np.concatenate([g[2] for g in sorted(gathered, key=lambda g: g[0])])
            krte = np.concatenate([g[3] for g in sorted(gathered, key=lambda g: g[0])])

            def rel(a, b):
                return float(np.max(np.abs(a - b) / np.maximum(np.abs(b), 1e-300)))
            errs = dict(Theta=rel(theta, single["Theta"]), k_rte=rel(krte, single["k_rte"]),
                        Beta=rel(mine["Beta"], single["Beta"]), t_rte=rel(mine["t_rte"], single["t_rte"]))
            good = all(v < tol for v in errs.values())
            print("world=%d dtype=%s overlapped=%s errs=%s replicas_identical(rank0 view)=%s -> %s" % (
                world, np.dtype(dtype).name, overlapped, {k_: "%.2e" % v for k_, v in errs.items()}, same,
                "OK" if good else "FAIL"), flush=True)
            ok = ok and good
        flag = torch.tensor([1 if same else 0], device=dev)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        if rank == 0 and int(flag.item()) != 1:
            print("item-side replicas differ across ranks -> FAIL", flush=True)
            ok = False
    dist.destroy_process_group()
    if rank == 0:
        print("MULTI_GPU_CHECK", "PASS" if ok else "FAIL", flush=True)
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
