"""User-sharded multi-GPU full-batch CAVI: one process per GPU, contiguous user ranges
balanced by nnz, a full replica of the item side on every rank, and ONE exchange per iteration: the
all-reduce (SUM) of the item-side partial sums (nI x ld reals) and of the k Theta column sums.
torch.distributed (NCCL over NVLink/NVSwitch) is the plumbing; all compute is the engine's kernels.

The prior `c` is added once, after the reduction (inside hpf_update_items), never per rank; every rank
then recomputes the identical item update from identical reduced data, so replicas stay bit-identical
without a broadcast.
"""
import numpy as np


def plan_user_shards(ix_u, nU, world):
    """Cut points (world+1 user ids, first 0, last nU) of contiguous user ranges holding ~equal nnz.
    `ix_u` may be a numpy array or a torch tensor (any device)."""
    if not isinstance(ix_u, np.ndarray) and hasattr(ix_u, "data_ptr"):   # torch tensor
        import torch
        deg = torch.bincount(ix_u.to(torch.int64), minlength=nU)
        csum = torch.cumsum(deg, 0)
        total = int(csum[-1].item()) if nU > 0 else 0
        targets = torch.tensor([total * r / world for r in range(1, world)], device=csum.device,
                               dtype=torch.float64)
        inner = (torch.searchsorted(csum.to(torch.float64), targets) + 1).clamp(max=nU).tolist() if world > 1 else []
    else:
        deg = np.bincount(np.asarray(ix_u, dtype=np.int64), minlength=nU)
        csum = np.cumsum(deg)
        total = int(csum[-1]) if nU > 0 else 0
        inner = [min(nU, int(np.searchsorted(csum, total * r / world) + 1)) for r in range(1, world)]
    cuts = [0] + [int(c) for c in inner] + [nU]
    for j in range(1, len(cuts)):            # monotone even for degenerate inputs
        cuts[j] = max(cuts[j], cuts[j - 1])
    return cuts


def shard_triples(ix_u, ix_i, Y, lo, hi):
    """Triples of users in [lo, hi) with user ids made local (numpy or torch, same type out)."""
    sel = (ix_u >= lo) & (ix_u < hi)
    return ix_u[sel] - lo, ix_i[sel], Y[sel]


class _RawCuda:
    def __init__(self, ptr, count, typestr):
        self.__cuda_array_interface__ = {"shape": (int(count),), "typestr": typestr,
                                         "data": (int(ptr), False), "version": 2}


def wrap_device_buffer(ptr, count, torch_dtype, device=None):
    """Zero-copy torch view of engine-owned device memory (for torch.distributed collectives)."""
    import torch
    typestr = {torch.float32: "<f4", torch.float64: "<f8"}[torch_dtype]
    dev = torch.device("cuda", torch.cuda.current_device() if device is None else device)
    return torch.as_tensor(_RawCuda(ptr, count, typestr), device=dev)


def engine_partial_tensors(engine, device=None):
    """(item_sums, theta_colsum) torch views of an Engine's reduction buffers."""
    import torch
    p1, n1, p2, n2 = engine.partials()
    real = torch.float32 if engine.real_bytes == 4 else torch.float64
    return wrap_device_buffer(p1, n1, real, device), wrap_device_buffer(p2, n2, torch.float64, device)


def run_sharded_iterations(engine, niter, partial_tensors=None, group=None, all_reduce=None):
    """`niter` full-batch iterations of one user shard.  `engine` is an hpfrec_b200.engine.Engine (or
    anything with sweep/update_users/update_items); the two tensors are all-reduced between the user
    and the item update.  With world size 1 (or no process group) this equals Engine.step_full."""
    import torch.distributed as dist
    if all_reduce is None:
        if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
            def all_reduce(t):
                dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
        else:
            def all_reduce(t):
                return None
    if partial_tensors is None:
        partial_tensors = engine_partial_tensors(engine)
    for _ in range(int(niter)):
        engine.sweep()
        engine.update_users()
        for t in partial_tensors:
            all_reduce(t)
        engine.update_items()


def run_sharded_iterations_overlapped(engine, niter, partial_tensors=None, group=None):
    """Same result as run_sharded_iterations, but the all-reduce of the item-side partial sums is
    issued asynchronously right after the item-major pass and overlaps the user-major pass and the
    user update (the NCCL kernel runs on NCCL's stream; `work.wait()` orders the item update after
    it on the engine's stream).  Needs an initialised process group and a real Engine."""
    import torch.distributed as dist
    if partial_tensors is None:
        partial_tensors = engine_partial_tensors(engine)
    t_items, t_theta = partial_tensors
    for _ in range(int(niter)):
        engine.sweep_side(0)
        w_items = dist.all_reduce(t_items, op=dist.ReduceOp.SUM, group=group, async_op=True)
        engine.sweep_side(1)
        engine.update_users()
        w_theta = dist.all_reduce(t_theta, op=dist.ReduceOp.SUM, group=group, async_op=True)
        w_items.wait()
        w_theta.wait()
        engine.update_items()


def attach_peers(engine, group=None):
    """Exchanges the CUDA-IPC handles of every rank's item-side buffers (all-gather over the process
    group) and maps them into `engine`, enabling run_sharded_iterations_peer."""
    import torch
    import torch.distributed as dist
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    mine = torch.frombuffer(bytearray(engine.peer_export()), dtype=torch.uint8).cuda()
    allh = [torch.empty_like(mine) for _ in range(world)]
    dist.all_gather(allh, mine, group=group)
    engine.peer_attach(rank, world, b"".join(bytes(t.cpu().numpy().tobytes()) for t in allh))


def run_sharded_iterations_peer(engine, niter, group=None, materialize_last=True):
    """User-sharded iterations with the item-side exchange fused into ONE kernel over NVLink peer
    memory (hpf_update_items_peer): each rank reduces its slice of item rows straight out of the other
    ranks' partial-sum buffers, updates it, and stores the result into every replica.  The only NCCL
    traffic left is two k-double all-reduces per iteration (Theta and Beta column sums), which double
    as the two cross-GPU barriers the kernel needs."""
    import torch
    import torch.distributed as dist
    dev = torch.cuda.current_device()
    _, _, p_theta, n_theta = engine.partials()
    p_beta, n_beta = engine.beta_colsum()
    t_theta = wrap_device_buffer(p_theta, n_theta, torch.float64, dev)
    t_beta = wrap_device_buffer(p_beta, n_beta, torch.float64, dev)
    for it in range(int(niter)):
        engine.sweep_side(0)
        engine.sweep_side(1)
        engine.update_users()
        dist.all_reduce(t_theta, op=dist.ReduceOp.SUM, group=group)
        engine.update_items_peer(materialize_last and it == niter - 1)
        dist.all_reduce(t_beta, op=dist.ReduceOp.SUM, group=group)
        engine.peer_finish()


class GraphedShardLoop:
    """One user-sharded iteration (kernels + NCCL collectives) captured ONCE into a CUDA graph and
    replayed, so that the per-iteration host cost is a single graph launch instead of ~7 Python ->
    C / c10d calls (at 8 GPUs the iteration is < 1 ms and the eager loop is host-bound).
    mode: "peer" (fused NVLink exchange, needs attach_peers), "overlap" or "plain" (NCCL all-reduce)."""

    def __init__(self, engine, mode="peer", group=None):
        import torch
        self.engine, self.mode, self.group = engine, mode, group
        self.stream = torch.cuda.Stream()
        self.graph = None
        self.launches_per_replay = 0   # engine kernels inside one captured iteration
        self.replayed_launches = 0     # kernels launched through graph replays (the engine cannot count those)
        engine.set_stream(self.stream)
        if mode == "peer":
            attach_peers(engine, group)

    def _one(self, materialize):
        if self.mode == "peer":
            run_sharded_iterations_peer(self.engine, 1, self.group, materialize_last=materialize)
        elif self.mode == "overlap":
            run_sharded_iterations_overlapped(self.engine, 1, group=self.group)
        else:
            run_sharded_iterations(self.engine, 1, group=self.group)

    def run(self, niter):
        """`niter` iterations; the last one runs eagerly so that shape/rate matrices are materialised."""
        import torch
        niter = int(niter)
        if niter <= 0:
            return
        cur = torch.cuda.current_stream()
        self.stream.wait_stream(cur)
        with torch.cuda.stream(self.stream):
            if self.graph is None and niter > 1:
                self._one(False)            # warm-up (NCCL channels, lazy allocations) before capture
                niter -= 1
                torch.cuda.synchronize()
                g = torch.cuda.CUDAGraph()
                before = self.engine.launch_count
                with torch.cuda.graph(g, stream=self.stream):
                    self._one(False)
                self.launches_per_replay = self.engine.launch_count - before
                self.graph = g
                self.engine.set_stream(self.stream)
            for _ in range(niter - 1):
                self.graph.replay()
                self.replayed_launches += self.launches_per_replay
            self._one(True)
        cur.wait_stream(self.stream)
