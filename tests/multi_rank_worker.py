"""Worker of tests/test_gpu_multi.py (one process per rank, launched with torch.distributed.run): runs
hpfrec_b200.dist.sharded_parity_check for every exchange mode and prints one JSON line per mode on rank 0.

HPF_TEST_SHARED_GPU=1: every rank uses cuda:0 (the driver's GPU test box has ONE GPU).  NCCL refuses two ranks
on one device, so the process group is gloo and the k-double all-reduces are staged through the host; the
item-side exchange kernel itself (update_items_peer_kernel: CUDA-IPC mapped peer buffers, P2P loads and
stores) runs exactly as it does across GPUs."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from hpfrec_b200 import dist as hdist  # noqa: E402


def main():
    shared = os.environ.get("HPF_TEST_SHARED_GPU", "0") == "1"
    local = 0 if shared else int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if shared:
        dist.init_process_group("gloo")
    else:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    ok = True
    modes = os.environ.get("HPF_TEST_MODES", "peer,overlap,plain" if shared else "nvls,symm,peer,overlap,plain").split(",")
    for mode in modes:
        fused = mode in ("peer", "nvls", "symm")
        # the fused exchange in its three schedules: reduce-scatter under the user-major pass, user update under the
        # item-major pass (the engine's factor buffers swap every iteration), nothing overlapped
        for overlap in ((True, "update", False) if fused else (None,)):
            for graph in ((False, True) if (not shared and mode in ("peer", "nvls")) else (False,)):
                its = 7 if graph else 3   # a replay holds two iterations: 7 = warm-up, capture, replays, odd remainder
                try:
                    res = hdist.sharded_parity_check(local, nU=20_000, nI=8_000, nnz=400_000, k=50, its=its, mode=mode,
                                                     graph=graph, overlap=overlap,
                                                     options={"overlap_update": 96} if overlap == "update" else None)
                except Exception as exc:  # a failed check must fail the test, not hang the other ranks
                    res = {"ok": False, "error": repr(exc)[:300], "what": "mode=%s graph=%s overlap=%s" % (mode, graph, overlap)}
                if dist.get_rank() == 0:
                    print("PARITY " + json.dumps(res), flush=True)
                ok = ok and bool(res.get("ok"))
    dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
