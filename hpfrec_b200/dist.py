"""User-sharded multi-GPU full-batch CAVI: one process per GPU, contiguous user ranges
balanced by nnz, a full replica of the item side on every rank, and ONE exchange per iteration: the
all-reduce (SUM) of the item-side partial sums (nI x ld reals) and of the k Theta column sums.
torch.distributed (NCCL over NVLink/NVSwitch) is the plumbing; all compute is the engine's kernels.

The prior `c` is added once, after the reduction (inside hpf_update_items), never per rank; every rank
then recomputes the identical item update from identical reduced data, so replicas stay bit-identical
without a broadcast.
"""
import os

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
                if t.is_cuda:
                    _all_reduce_device(t, group)
                else:
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
        w_items = _all_reduce_device(t_items, group, async_op=True)
        engine.sweep_side(1)
        engine.update_users()
        w_theta = _all_reduce_device(t_theta, group, async_op=True)
        for w in (w_items, w_theta):
            if w is not None:
                w.wait()
        engine.update_items()


def _all_reduce_device(t, group=None, async_op=False):
    """SUM all-reduce of a device tensor.  With NCCL it is the collective itself; with a CPU backend (gloo:
    the 2-ranks-on-one-GPU test set-up, where NCCL refuses duplicate devices) it is staged through the host."""
    import torch.distributed as dist
    if dist.get_backend(group) == "nccl":
        return dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group, async_op=async_op)
    host = t.cpu()
    dist.all_reduce(host, op=dist.ReduceOp.SUM, group=group)
    t.copy_(host)
    return None


def _all_gather_bytes(payload, group=None):
    """All-gather of one equal-length bytes object per rank (CPU or NCCL backend)."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group)
    mine = torch.frombuffer(bytearray(payload), dtype=torch.uint8)
    if dist.get_backend(group) == "nccl":
        mine = mine.cuda()
    allh = [torch.empty_like(mine) for _ in range(world)]
    dist.all_gather(allh, mine, group=group)
    return [bytes(t.cpu().numpy().tobytes()) for t in allh]


def attach_peers(engine, group=None):
    """Exchanges the CUDA-IPC handles of every rank's item-side buffers (all-gather over the process
    group) and maps them into `engine`, enabling run_sharded_iterations_peer."""
    import torch
    import torch.distributed as dist
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    engine.peer_attach(rank, world, b"".join(_all_gather_bytes(engine.peer_export(), group)))


def attach_symmetric(engine, group=None, multicast=True):
    """Moves the engine's five item-side buffers into ONE symmetric allocation (torch.distributed._symmetric_memory:
    every rank allocates the same size, the rendezvous maps every rank's copy into this process and, on NVSwitch
    systems, also creates a multicast mapping), then hands the peer and multicast pointers to the engine.  Must run
    before hpf_load_state.  Returns (handle, tensor, used_multicast); keep both alive as long as the engine."""
    import torch
    import torch.distributed as dist
    import torch.distributed._symmetric_memory as symm_mem
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    sizes = engine.item_buffer_bytes()
    offs, total = [], 0
    for nbytes in sizes:
        offs.append(total)
        total += (nbytes + 4095) // 4096 * 4096
    dev = torch.device("cuda", torch.cuda.current_device())
    t = symm_mem.empty(total, dtype=torch.uint8, device=dev)
    hdl = symm_mem.rendezvous(t, group if group is not None else dist.group.WORLD)
    bases = [int(p) for p in hdl.buffer_ptrs]
    if bases[rank] != t.data_ptr():
        raise RuntimeError("symmetric memory: this rank's mapped pointer is not the tensor's own")
    mc_base = int(getattr(hdl, "multicast_ptr", 0) or 0) if multicast else 0
    engine.adopt_item_buffers([t.data_ptr() + o for o in offs])
    peer_ptrs = [[bases[p] + o for o in offs] for p in range(world)]
    engine.peer_attach_ptrs(rank, world, peer_ptrs, [mc_base + o for o in offs] if mc_base else None)
    return hdl, t, bool(mc_base)


class _PhaseTimer:
    """HPF_PHASES=1: CUDA events between the phases of the eager sharded loop (main stream), summed over iterations
    and printed by every rank as one `PHASES {...}` JSON line."""

    totals = {}
    iterations = 0

    def __init__(self):
        self.events = []

    def mark(self, name):
        import torch
        ev = torch.cuda.Event(enable_timing=True)
        ev.record()
        self.events.append((name, ev))

    def report(self):
        import json
        import torch
        import torch.distributed as dist
        torch.cuda.synchronize()
        cls = _PhaseTimer
        for (n0, e0), (n1, e1) in zip(self.events[:-1], self.events[1:]):
            if n1 == "start":
                continue
            cls.totals[n1] = cls.totals.get(n1, 0.0) + e0.elapsed_time(e1)
        cls.iterations += sum(1 for n, _ in self.events if n == "start")
        per = {k: round(v / max(cls.iterations, 1), 4) for k, v in cls.totals.items()}
        per["sum"] = round(sum(per.values()), 4)
        print("PHASES " + json.dumps({"rank": dist.get_rank() if dist.is_initialized() else 0, "iterations": cls.iterations,
                                      "ms_per_iteration": per}), flush=True)


_side_streams = {}


def _side_stream(dev):
    """One high-priority side stream per device for the overlapped reduce-scatter."""
    import torch
    if dev not in _side_streams:
        _side_streams[dev] = (torch.cuda.Stream(device=dev, priority=-1), torch.zeros(1, dtype=torch.float32, device=dev))
    return _side_streams[dev]


class SymmSync:
    """Cross-GPU synchronisation of the sharded loop over symmetric memory instead of NCCL: device-side barriers on
    the allocation's signal pads (one tiny kernel, a few microseconds, against ~20 us for a k-double NCCL all-reduce)
    and the two k-double column-sum exchanges as "publish my partial, barrier, add everybody's partials in rank
    order" (every rank adds the same numbers in the same order: bit-identical sums on all ranks)."""

    TIMEOUT_MS = 60000   # a barrier that is not met traps instead of hanging the GPU

    def __init__(self, n, group=None):
        import torch
        import torch.distributed as dist
        import torch.distributed._symmetric_memory as symm_mem
        self.world, self.rank = dist.get_world_size(group), dist.get_rank(group)
        dev = torch.device("cuda", torch.cuda.current_device())
        self.n = int(n)
        self.buf = symm_mem.empty(2 * self.n, dtype=torch.float64, device=dev)
        self.hdl = symm_mem.rendezvous(self.buf, group if group is not None else dist.group.WORLD)
        self.buf.zero_()
        self.views = [self.hdl.get_buffer(p, (2, self.n), torch.float64) for p in range(self.world)]
        torch.cuda.synchronize()
        self.hdl.barrier(channel=3, timeout_ms=self.TIMEOUT_MS)
        torch.cuda.synchronize()

    def barrier(self, channel):
        self.hdl.barrier(channel=channel, timeout_ms=self.TIMEOUT_MS)

    def all_reduce(self, t, slot, channel):
        """t (n doubles, device) <- sum over ranks of t.  The slot is rewritten only after every rank has passed at
        least one later barrier, i.e. after it has read this one."""
        import torch
        self.views[self.rank][slot].copy_(t)
        self.barrier(channel)
        torch.sum(torch.stack([v[slot] for v in self.views]), dim=0, out=t)


_side_streams = {}


def _side_stream(dev):
    """One high-priority side stream per device for the overlapped reduce-scatter."""
    import torch
    if dev not in _side_streams:
        _side_streams[dev] = (torch.cuda.Stream(device=dev, priority=-1), torch.zeros(1, dtype=torch.float32, device=dev))
    return _side_streams[dev]


def run_sharded_iterations_peer(engine, niter, group=None, materialize_last=True, overlap=True, sync=None):
    """User-sharded iterations with the item-side exchange over NVLink peer / NVSwitch multicast memory: each rank
    reduces its slice of item rows straight out of all ranks' partial-sum buffers, updates it, and stores the result
    into every replica.  Besides that kernel only k-double sums cross GPUs (Theta and Beta column sums), and they
    double as the cross-GPU barriers the kernels need: NCCL all-reduces, or with `sync` (a SymmSync) device-side
    barriers and peer reads over symmetric memory.

    overlap="update": user-major pass first, then the item-major pass with the USER UPDATE under it on the engine's
    second stream (hpf_item_pass_with_user_update), then the single fused exchange kernel.  The engine's two user-factor
    buffers swap roles every iteration, so a captured loop must hold an even number of iterations.
    overlap=True splits the exchange: the reduce-scatter half (hpf_reduce_items_peer) starts right after the
    item-major pass on a second stream -- behind one extra barrier that makes every rank's partial sums final -- and
    runs UNDER the user-major pass and the user update; the update + broadcast half follows on the main stream.
    overlap=False is the single fused kernel after both passes."""
    import torch
    import torch.distributed as dist
    dev = torch.cuda.current_device()
    _, _, p_theta, n_theta = engine.partials()
    p_beta, n_beta = engine.beta_colsum()
    t_theta = wrap_device_buffer(p_theta, n_theta, torch.float64, dev)
    t_beta = wrap_device_buffer(p_beta, n_beta, torch.float64, dev)
    main = torch.cuda.current_stream()
    update_under_pass = overlap == "update"
    if update_under_pass:
        overlap = False
    if overlap:
        side, t_bar = _side_stream(dev)
    phases = _PhaseTimer() if (os.environ.get("HPF_PHASES") == "1" and not torch.cuda.is_current_stream_capturing()) else None
    for it in range(int(niter)):
        last = materialize_last and it == niter - 1
        if phases:
            phases.mark("start")
        if update_under_pass:
            # user-major pass, then the item-major pass with the user update under it (engine's second stream); the
            # Theta sums below are also the barrier that makes every rank's item-side partial sums final
            engine.sweep_side(1)
            if phases:
                phases.mark("user-major pass")
            engine.item_pass_with_user_update(last)
            if phases:
                phases.mark("item-major pass || user update")
            if sync is not None:
                sync.all_reduce(t_theta, 0, 1)
            else:
                _all_reduce_device(t_theta, group)
            if phases:
                phases.mark("Theta column sums (barrier)")
            engine.update_items_peer(last)
            if phases:
                phases.mark("item update + exchange")
            if sync is not None:
                sync.all_reduce(t_beta, 1, 2)
            else:
                _all_reduce_device(t_beta, group)
            if phases:
                phases.mark("Beta column sums (barrier)")
            engine.peer_finish()
            if phases:
                phases.mark("re-zero item sums")
            continue
        engine.sweep_side(0)
        if phases:
            phases.mark("item-major pass")
        if overlap:
            side.wait_stream(main)
            with torch.cuda.stream(side):
                if sync is not None:                        # every rank's item-major pass is done
                    sync.barrier(0)
                else:
                    _all_reduce_device(t_bar, group)
                engine.reduce_items_peer(side.cuda_stream)
        engine.sweep_side(1)
        if phases:
            phases.mark("user-major pass")
        engine.update_users(last)
        if phases:
            phases.mark("user update")
        if sync is not None:
            sync.all_reduce(t_theta, 0, 1)
        else:
            _all_reduce_device(t_theta, group)
        if phases:
            phases.mark("Theta column sums (barrier)")
        if overlap:
            main.wait_stream(side)
        if phases:
            phases.mark("wait for the reduce-scatter")
        engine.update_items_peer(last)
        if phases:
            phases.mark("item update + broadcast")
        if sync is not None:
            sync.all_reduce(t_beta, 1, 2)
        else:
            _all_reduce_device(t_beta, group)
        if phases:
            phases.mark("Beta column sums (barrier)")
        engine.peer_finish()
        if phases:
            phases.mark("re-zero item sums")
    if phases:
        phases.report()


def resolve_overlap(n_users_local):
    """The exchange schedule of the fused modes for a shard of `n_users_local` users: HPF_EXCHANGE_OVERLAP =
    auto | 1 (reduce-scatter under the user-major pass) | update (user update under the item-major pass) | 0.
    auto, measured (profiles/r02_bench_C5_N8_update.json, r02_bench_N8_schedule_*.json, r02_bench_N2_overlap_*.json):
    hiding the user update wins on C5 at 8 GPUs (252 vs 242-244 it/s) and on H at 8 GPUs (1462 vs 1441 on the same
    box) and ties on H at 2 GPUs (1.446 vs 1.441 ms); tiny shards keep the reduce-scatter overlap (the update is too
    short to be worth a second stream there)."""
    env = os.environ.get("HPF_EXCHANGE_OVERLAP", "auto")
    if env == "auto":
        return "update" if n_users_local >= 100_000 else True
    return "update" if env == "update" else env != "0"


class ShardedLoop:
    """Iteration driver of one user shard: picks the item-side exchange and, optionally, replays one captured
    iteration as a CUDA graph (kernels + collectives) so that the per-iteration host cost is a single graph
    launch instead of ~9 Python -> C / c10d calls.

    mode: "nvls"    fused reduce-scatter + item update + all-gather kernel over NVSwitch multicast memory:
                    multimem.ld_reduce sums every rank's partial sums inside the switch, multimem.st broadcasts
                    the updated rows (symmetric memory; falls back to "symm" when the system has no multicast)
          "symm"    the same kernel with peer loads / stores over symmetric memory
          "peer"    the same kernel over CUDA-IPC mapped buffers
          "overlap" NCCL all-reduce of the item-side partial sums, overlapping the user-major pass
          "plain"   NCCL all-reduce after both passes
          "auto"    "nvls" with NCCL, else "peer"
    Construct it BEFORE hpf_load_state (the fused modes map the engine's item-side buffers into every rank)."""

    ITERATIONS_PER_REPLAY = 2   # with overlap="update" (the engine's user-factor buffers swap every iteration)

    def __init__(self, engine, mode="auto", group=None, graph=False, overlap=None):
        import torch
        import torch.distributed as dist
        self.engine, self.group = engine, group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        if mode == "auto":
            mode = "nvls" if (dist.is_initialized() and dist.get_backend(group) == "nccl") else "peer"
        if self.world == 1:
            mode = "single"
        if mode not in ("nvls", "symm", "peer", "overlap", "plain", "single"):
            raise ValueError("unknown exchange mode %r" % (mode,))
        self._symm = None
        if mode in ("nvls", "symm"):
            try:
                hdl, t, used_mc = attach_symmetric(engine, group, multicast=(mode == "nvls"))
                self._symm = (hdl, t)
                mode = "nvls" if used_mc else "symm"
            except Exception as exc:   # no symmetric-memory support on this system: CUDA IPC does the same job
                import warnings
                warnings.warn("symmetric memory unavailable (%r): using the CUDA-IPC peer exchange" % (exc,))
                mode = "peer"
        self.mode = mode
        #: reduce-scatter half of the fused exchange on a second stream, under the user-major pass (HPF_EXCHANGE_OVERLAP=0: off)
        if overlap is None:
            overlap = resolve_overlap(engine.nU)
        self.overlap = overlap
        #: barriers and k-double sums over symmetric memory instead of NCCL (needs the symmetric allocation; HPF_SYNC=nccl: off)
        self.sync = None
        if self._symm is not None and os.environ.get("HPF_SYNC", "symm") == "symm":
            try:
                self.sync = SymmSync(engine.k, group)
            except Exception as exc:
                import warnings
                warnings.warn("symmetric-memory barriers unavailable (%r): using NCCL all-reduces" % (exc,))
        self.use_graph = bool(graph) and self.world > 1
        self.stream = torch.cuda.Stream() if self.use_graph else torch.cuda.current_stream()
        self.graph = None
        self._swaps = 0
        self.launches_per_replay = 0   # engine kernels inside one captured replay
        self.replayed_launches = 0     # kernels launched through graph replays (the engine cannot count those)
        if self.use_graph:
            engine.set_stream(self.stream)
        if engine.describe().get("robust") == "1" and self.world > 1:
            raise ValueError("robust mode (tiny shape priors) is not available for sharded runs")
        if mode == "peer":
            attach_peers(engine, group)

    def describe(self):
        """exchange mode + how the ranks synchronise, for telemetry"""
        if self.mode in ("nvls", "symm", "peer"):
            how = {True: ", reduce-scatter overlapped", False: "", "update": ", user update under the item-major pass"}[self.overlap]
            return "%s%s, %s barriers" % (self.mode, how,
                                          "symmetric-memory" if self.sync is not None else "NCCL")
        return self.mode

    def _iterations(self, n, materialize_last):
        if self.overlap == "update":
            self._swaps = (self._swaps + int(n)) % 2   # user-factor buffer roles relative to the captured graph
        if self.mode == "single":
            self.engine.step_full(n)
        elif self.mode in ("peer", "nvls", "symm"):
            run_sharded_iterations_peer(self.engine, n, self.group, materialize_last=materialize_last, overlap=self.overlap,
                                        sync=self.sync)
        elif self.mode == "overlap":
            run_sharded_iterations_overlapped(self.engine, n, group=self.group)
        else:
            run_sharded_iterations(self.engine, n, group=self.group)

    def run(self, niter):
        """`niter` iterations; with a graph, the last one runs eagerly so that shape/rate matrices are materialised."""
        import torch
        niter = int(niter)
        if niter <= 0:
            return
        if not self.use_graph:
            self._iterations(niter, True)
            return
        cur = torch.cuda.current_stream()
        self.stream.wait_stream(cur)
        with torch.cuda.stream(self.stream):
            # a captured replay holds TWO iterations: the engine's user-factor buffers swap roles every iteration when the
            # user update runs under the item-major pass, and a replay must leave them as it found them
            per = self.ITERATIONS_PER_REPLAY if self.overlap == "update" else 1
            if self.graph is None and niter > per + 1:
                self._iterations(per, False)   # warm-up (NCCL channels, lazy allocations, second stream) before capture
                niter -= per
                self.stream.synchronize()
                self._swaps = 0                # the graph's buffer roles are the ones that hold right now
                g = torch.cuda.CUDAGraph()
                before = self.engine.launch_count
                with torch.cuda.graph(g, stream=self.stream, capture_error_mode="thread_local"):
                    self._iterations(per, False)
                self.launches_per_replay = self.engine.launch_count - before
                self.graph = g
            while self.graph is not None and niter > per:
                if self._swaps:   # an odd number of eager iterations since the capture: the graph's buffer roles do not hold
                    self._iterations(1, False)
                    niter -= 1
                    continue
                self.graph.replay()
                self.replayed_launches += self.launches_per_replay
                niter -= per
            if niter > 1:
                self._iterations(niter - 1, False)
            if niter >= 1:
                self._iterations(1, True)
        cur.wait_stream(self.stream)

    def close(self):
        """Call AFTER the engine is closed when symmetric memory is in use (the engine's item buffers live in it)."""
        self.graph = None
        self.sync = None
        self._symm = None


def sharded_parity_check(local_device, options=None, nU=60_000, nI=25_000, nnz=1_500_000, k=50, its=3, mode=None,
                         graph=False, group=None, overlap=None):
    """Driver-visible proof of the multi-GPU data plane (run by bench.py before it times anything at N>1):
    `its` fp64 iterations of a down-scaled problem, user-sharded over all ranks with the SAME exchange the
    bench is about to time, against one engine on rank 0.  Checks (i) every state array <= 1e-10 relative,
    (ii) the item-side replicas bit-identical across ranks.  Returns a dict for the JSON line; raises on failure."""
    import os
    import torch
    import torch.distributed as dist
    import bench
    from .engine import Engine
    from .loops import CudaLoops
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    dev = torch.device("cuda", local_device)
    mode = mode or os.environ.get("HPF_MULTI", "auto")
    u, i, y = bench.synth_coo_torch(nU, nI, nnz, dev, seed=7)
    y = y.to(torch.float64)
    loops = CudaLoops(False, device=local_device)
    Gs, Gr, Ls, Lr, kr, tr = loops.initialize_parameters(np.empty((nU, k)), np.empty((nI, k)), 123, 0.3, 0.3, 1.0, 0.3, 0.3, 1.0)
    cuts = plan_user_shards(u, nU, world)
    lo, hi = cuts[rank], cuts[rank + 1]
    lu, li, ly = (t.contiguous() for t in shard_triples(u, i, y, lo, hi))
    eng = Engine(hi - lo, nI, k, 8, local_device)
    for name, val in (options or {}).items():
        eng.set_option(name, val)
    loop = ShardedLoop(eng, mode=mode, graph=graph, group=group, overlap=overlap)
    eng.load_state(np.ascontiguousarray(Gs[lo:hi]), np.ascontiguousarray(Gr[lo:hi]), Ls, Lr, np.ascontiguousarray(kr[lo:hi]), tr)
    eng.load_coo(lu, li, ly)
    loop.run(its)
    torch.cuda.synchronize()
    mine = eng.export_all()
    how = loop.describe()
    eng.close()
    loop.close()
    cdev = dev if dist.get_backend(group) == "nccl" else torch.device("cpu")
    beta = torch.from_numpy(mine["Beta"]).to(cdev)
    ref_beta = beta.clone()
    dist.broadcast(ref_beta, 0, group=group)
    same = torch.tensor([1 if torch.equal(beta, ref_beta) else 0], device=cdev)
    dist.all_reduce(same, op=dist.ReduceOp.MIN, group=group)
    gathered = [None] * world
    dist.all_gather_object(gathered, (lo, mine["Theta"], mine["k_rte"], mine["Gamma_shp"]), group=group)
    result = None
    if rank == 0:
        e1 = Engine(nU, nI, k, 8, local_device)
        e1.load_state(Gs, Gr, Ls, Lr, kr, tr)
        e1.load_coo(u.contiguous(), i.contiguous(), y.contiguous())
        e1.step_full(its)
        single = e1.export_all()
        e1.close()
        parts = sorted(gathered, key=lambda t: t[0])

        def rel(x, ref):
            return float(np.max(np.abs(x - ref) / np.maximum(np.abs(ref), 1e-300)))
        errs = dict(Theta=rel(np.concatenate([p[1] for p in parts]), single["Theta"]),
                    k_rte=rel(np.concatenate([p[2] for p in parts]), single["k_rte"]),
                    Gamma_shp=rel(np.concatenate([p[3] for p in parts]), single["Gamma_shp"]),
                    Beta=rel(mine["Beta"], single["Beta"]), Lambda_shp=rel(mine["Lambda_shp"], single["Lambda_shp"]),
                    Lambda_rte=rel(mine["Lambda_rte"], single["Lambda_rte"]), t_rte=rel(mine["t_rte"], single["t_rte"]))
        result = {"what": "%d fp64 iterations of %dx%dx%d k=%d sharded over %d GPUs (exchange=%s%s) vs one engine"
                          % (its, nU, nI, nnz, k, world, how, ", graph replay" if graph else ""),
                  "max_rel_err": max(errs.values()), "errors": errs, "tolerance": 1e-10,
                  "item_replicas_bit_identical": bool(int(same.item()) == 1)}
        result["ok"] = bool(result["max_rel_err"] < 1e-10 and result["item_replicas_bit_identical"])
    box = [result]
    dist.broadcast_object_list(box, 0, group=group)
    if not box[0]["ok"]:
        raise RuntimeError("multi-GPU parity check FAILED: %r" % (box[0],))
    return box[0]
