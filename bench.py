#!/usr/bin/env python3
"""bench.py -- full-batch CAVI iterations/s on the MillionSong-shaped synthetic (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W [--config H|C2|C3|H_f64|C4|C5] [--impl reference]

A "step" is one complete full-batch CAVI iteration (the reference's loop body, cython_loops.pxi:230-259)
over the whole nnz list -- for --config C4 one SVI epoch (cython_loops.pxi:262-377).  Default workload = H
of SURVEY.md §8: 1M x 380K users x items, 48M nnz, k=50, fp32.  For N>1 the nnz list is sharded by user
range (balanced by nnz) with the item side replicated and one exchange per iteration; the problem size is
fixed, so scaling is "strong".

Rank 0 prints ONE JSON line (see the contract in the task description): value = iterations/s from
device-resident inputs (CUDA events on the engine's stream, barrier + synchronize on both sides, max
over ranks); `e2e` = the same iterations through the C ABI with HOST (pinned) buffers: upload of state
and triples, index build, K iterations, download of the eight result arrays, all inside the timed
region; `roofline` = SURVEY §8(d)'s algorithmic bytes of one iteration / device time per iteration against
MEASURED_PEAKS.json; `cpu_baseline` = the compiled, unmodified reference (oracle/_ref) timed on this
box's host cores on a bounded sample.  `--impl reference` times that reference on the FULL configuration.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "full_batch_cavi_iterations_per_s"
UNIT = "iterations/s"

# BASELINE.json configs (SURVEY §8 preamble): shapes, row length, arithmetic type, minibatch sizes
CONFIGS = {
    "H": dict(nusers=1_000_000, nitems=380_000, nnz=48_000_000, k=50, dtype="f32"),
    "C2": dict(nusers=1_000_000, nitems=380_000, nnz=48_000_000, k=30, dtype="f32"),
    "C3": dict(nusers=1_000_000, nitems=380_000, nnz=48_000_000, k=128, dtype="f32"),
    "H_f64": dict(nusers=1_000_000, nitems=380_000, nnz=48_000_000, k=50, dtype="f64"),
    "C4": dict(nusers=1_000_000, nitems=380_000, nnz=48_000_000, k=50, dtype="f32", users_per_batch=50_000,
               items_per_batch=20_000),
    "C5": dict(nusers=10_000_000, nitems=1_000_000, nnz=500_000_000, k=50, dtype="f32"),
}


# -----------------------------------------------------------------------------------------------------
def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default=None, choices=sorted(CONFIGS), help="a BASELINE.json configuration (default H)")
    ap.add_argument("--nusers", type=int, default=None)
    ap.add_argument("--nitems", type=int, default=None)
    ap.add_argument("--nnz", type=int, default=None)
    ap.add_argument("--k", type=int, default=None)
    ap.add_argument("--dtype", default=None, choices=["f32", "f64"])
    ap.add_argument("--users-per-batch", type=int, default=None)
    ap.add_argument("--items-per-batch", type=int, default=None)
    ap.add_argument("--alpha", type=float, default=0.6, help="Zipf exponent of item popularity")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-parity-check", action="store_true")
    ap.add_argument("--cpu-sample-div", type=int, default=8, help="cpu_baseline sample = workload / this")
    ap.add_argument("--ref-sample-div", type=int, default=1,
                    help="--impl reference: 1 = the full configuration (default); >1 times workload / this and says so")
    ap.add_argument("--init", default=None, choices=["host", "device"],
                    help="random start: host = the reference's numpy MT19937 stream (default), device = torch generator "
                         "(default for C5, whose host start is 4.4 GB per rank)")
    ap.add_argument("--option", action="append", default=[], help="engine option name=value")
    a = ap.parse_args()
    base = dict(CONFIGS[a.config or "H"])
    for key in ("nusers", "nitems", "nnz", "k", "dtype"):
        if getattr(a, key) is None:
            setattr(a, key, base[key])
    if a.users_per_batch is None:
        a.users_per_batch = base.get("users_per_batch", 0)
    if a.items_per_batch is None:
        a.items_per_batch = base.get("items_per_batch", 0)
    a.config_name = a.config or ("H" if all(getattr(a, k_) == CONFIGS["H"][k_] for k_ in CONFIGS["H"]) else "custom")
    if a.init is None:
        a.init = "device" if a.nusers * a.k > 200_000_000 else "host"
    return a


def workload_name(a, div=1):
    nU, nI, nnz = a.nusers // div, a.nitems // div, a.nnz // div
    kind = "full-batch CAVI" if not (a.users_per_batch or a.items_per_batch) else \
        "SVI epochs (%d users / %d items per batch)" % (a.users_per_batch, a.items_per_batch)
    name = "%s: %dx%d users x items, %d nnz, k=%d, %s %s (synthetic: lognormal users, Zipf(%.1f) items)" % (
        a.config_name, nU, nI, nnz, a.k, "fp32" if a.dtype == "f32" else "fp64", kind, a.alpha)
    if div != 1:
        name += " [SAMPLE: workload / %d]" % div
    return name


# -----------------------------------------------------------------------------------------------------
def synth_coo_torch(nU, nI, nnz, device, seed=42, alpha=0.6):
    """Device-side version of oracle.hpf_oracle.synth_coo's recipe: same distributions,
    torch generator (fixed seed, identical on every rank).  Returns int64 u, i and float32 y."""
    import torch
    g = torch.Generator(device=device)
    g.manual_seed(seed)
    pu = torch.exp(torch.randn(nU, generator=g, device=device, dtype=torch.float64))
    pi = torch.arange(1, nI + 1, device=device, dtype=torch.float64) ** (-alpha)
    pi = pi[torch.randperm(nI, generator=g, device=device)]
    cu = torch.cumsum(pu / pu.sum(), 0)
    ci = torch.cumsum(pi / pi.sum(), 0)
    keys = torch.empty(0, dtype=torch.int64, device=device)
    while keys.numel() < nnz:
        need = int((nnz - keys.numel()) * 1.25) + 16
        u = torch.searchsorted(cu, torch.rand(need, generator=g, device=device, dtype=torch.float64)).clamp_(max=nU - 1)
        i = torch.searchsorted(ci, torch.rand(need, generator=g, device=device, dtype=torch.float64)).clamp_(max=nI - 1)
        keys = torch.unique(torch.cat([keys, u * nI + i]))
        del u, i
    keys = keys[torch.randperm(keys.numel(), generator=g, device=device)[:nnz]]
    r = torch.rand(nnz, generator=g, device=device, dtype=torch.float64)
    y = torch.clamp(1 + torch.floor((1 - r) ** (-1 / 1.5) - 1), max=1e4).to(torch.float32)
    return keys // nI, keys % nI, y


def algorithmic_bytes(nU, nI, nnz, k, s):
    """stream the triples once + read and write each of the four state matrices once."""
    return nnz * (4 + 4 + s) + 4 * (nU + nI) * k * s


class ClockSampler:
    QUERY = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
             "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index=0):
        self.rows, self.proc, self.gpu = [], None, gpu_index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + self.QUERY,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), line.strip()))

    def stop(self, t0, t1):
        if self.proc is None:
            return None
        time.sleep(0.15)
        self.proc.terminate()
        sm, smax, reasons, power = [], [], set(), []
        for ts, line in self.rows:
            if ts < t0 - 0.05 or ts > t1 + 0.15:
                continue
            f = [x.strip() for x in line.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                smax.append(float(f[2]))
                power.append(float(f[3]))
            except ValueError:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(smax)), "reasons": sorted(reasons),
                "samples": len(sm), "power_w_max": float(max(power))}


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def cpu_info():
    cores = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    model = ""
    try:
        for line in open("/proc/cpuinfo"):
            if line.startswith("model name"):
                model = line.split(":", 1)[1].strip()
                break
    except Exception:
        pass
    return cores, model


def host_triples(a, div):
    """The workload's triples on the host.  Generated on the GPU when there is one (identical to the data of
    the GPU arm), else with the numpy generator of the oracle (same recipe, slower)."""
    nU, nI, nnz = max(64, a.nusers // div), max(64, a.nitems // div), max(1024, a.nnz // div)
    try:
        import torch
        if torch.cuda.is_available():
            u, i, y = synth_coo_torch(nU, nI, nnz, torch.device("cuda", 0), seed=42, alpha=a.alpha)
            out = (u.cpu().numpy(), i.cpu().numpy(), y.cpu().numpy().astype(np.float64))
            del u, i, y
            torch.cuda.empty_cache()
            return nU, nI, nnz, out, "torch generator on the GPU (the GPU arm's data)"
    except Exception:
        pass
    from oracle import hpf_oracle as O
    return nU, nI, nnz, O.synth_coo(nU, nI, nnz, seed=42, alpha=a.alpha), "numpy generator (oracle.synth_coo)"


# -----------------------------------------------------------------------------------------------------
def cpu_reference_rate(a, warm_iters, timed_iters, div):
    """Times the compiled UNMODIFIED reference (oracle/_ref: hpfrec.cython_loops_{float,double}.fit_hpf, all host
    threads, deterministic scatter) on the workload / div.  Two calls: fit_hpf(maxiter=warm_iters) and
    fit_hpf(maxiter=warm_iters+timed_iters); both pay the same initialisation and phi allocation, so their
    difference is `timed_iters` iterations.  Returns dict or None if oracle/_ref is absent."""
    from oracle import ref_loader as R
    use_float = a.dtype == "f32"
    mod = R.load(use_float)
    if mod is None:
        return None
    nU, nI, nnz, (u, i, y), gen = host_triples(a, div)
    cores, model = cpu_info()
    dt = np.float32 if use_float else np.float64
    y = y.astype(dt)
    upb, ipb = a.users_per_batch // div, a.items_per_batch // div
    st_ix_u = None
    ncores = cores
    if upb or ipb:   # SVI: data sorted by user + CSR pointer (init:516-521); only ncores=1 is deterministic
        order = np.argsort(u, kind="stable")
        u, i, y = u[order], i[order], y[order]
        st_ix_u = np.concatenate([[0], np.cumsum(np.bincount(u, minlength=nU))]).astype(np.uint64)

    def run(iters):
        t0 = time.time()
        R.ref_fit_hpf(mod, y, u, i, nU, nI, a.k, iters, seed=123, ncores=ncores, par_sh=0, users_per_batch=upb,
                      items_per_batch=ipb, st_ix_u=st_ix_u)
        return time.time() - t0

    warm_iters = max(1, warm_iters)
    t_warm = run(warm_iters)
    t_all = run(warm_iters + timed_iters)
    s_per_iter = max(1e-9, (t_all - t_warm) / timed_iters)
    return {"s_per_iter": s_per_iter, "nnz_per_s": nnz / s_per_iter, "cores": cores, "cpu_model": model,
            "seconds_warm_call": t_warm, "seconds_timed_call": t_all, "nU": nU, "nI": nI, "nnz": nnz,
            "data": gen,
            "sample": "%dx%d, %d nnz (workload/%d), k=%d, %s, (t[maxiter=%d]-t[maxiter=%d])/%d, %d threads"
                      % (nU, nI, nnz, div, a.k, "fp32" if use_float else "fp64", warm_iters + timed_iters, warm_iters,
                         timed_iters, ncores)}


def run_reference(a):
    """The reference arm: the reference's own CPU implementation, its stock entry point (fit_hpf, pxi:147), all
    host threads, on the SAME configuration as the GPU arm (unless --ref-sample-div says otherwise).  One iteration
    costs ~13 s at the headline size, so the warm-up is ONE call with maxiter=1 (it pays the same initialisation and
    phi allocation as the timed call, and first-touches the pages); the timed call runs K+1 iterations and
    ms_per_step = (t[K+1] - t[1]) / K: exactly K timed iterations, all of them executed."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    div = max(1, a.ref_sample_div)
    res = cpu_reference_rate(a, 1, a.steps, div)
    if res is None:
        print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref not built on this box"}))
        return
    v = 1.0 / res["s_per_iter"]
    note = "the full configuration" if div == 1 else "a 1/%d sample of the configuration (labelled in config.workload)" % div
    line = {"impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": a.gpus,
            "steps": a.steps, "warmup": a.warmup, "ms_per_step": 1000.0 * res["s_per_iter"], "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": a.dtype, "data": "synthetic",
            "config": {"workload": workload_name(a, div), "what": "compiled unmodified reference on " + note,
                       "data_generator": res["data"]},
            "nnz_per_s": res["nnz_per_s"],
            "cpu_baseline": {"value": v, "unit": UNIT, "cores": res["cores"], "kind": "reference",
                             "sample": res["sample"], "cpu_model": res["cpu_model"],
                             "seconds_warm_call": res["seconds_warm_call"], "seconds_timed_call": res["seconds_timed_call"],
                             "build": "oracle/build_ref.py: gcc -O2 -fopenmp -march=x86-64-v3 (the reference's setup.py uses "
                                      "-march=native -flto); the serial scatter bounds it, 16 and 32 threads give the same rate"},
            "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line))


# -----------------------------------------------------------------------------------------------------
def initial_state(a, loops, lo, hi, dev):
    """(Gamma_shp, Gamma_rte, Lambda_shp, Lambda_rte, k_rte, t_rte) for users [lo, hi) and all items."""
    import torch
    nU, nI, k = a.nusers, a.nitems, a.k
    npdt = np.float32 if a.dtype == "f32" else np.float64
    if a.init == "host":
        Theta = np.empty((nU, k), npdt)
        Beta = np.empty((nI, k), npdt)
        Gs, Gr, Ls, Lr, kr, tr = loops.initialize_parameters(Theta, Beta, 123, 0.3, 0.3, 1.0, 0.3, 0.3, 1.0)
        del Theta, Beta
        return [np.ascontiguousarray(x[lo:hi]) for x in (Gs, Gr)] + [Ls, Lr, np.ascontiguousarray(kr[lo:hi]), tr]
    # device start (same distribution as pxi:134-138: prior + 0.01 U; NOT the reference's bit stream)
    tdt = torch.float32 if a.dtype == "f32" else torch.float64
    g = torch.Generator(device=dev)
    g.manual_seed(123)
    Lr = 0.3 + 0.01 * torch.rand((nI, k), generator=g, device=dev, dtype=tdt)
    Ls = 0.3 + 0.01 * torch.rand((nI, k), generator=g, device=dev, dtype=tdt)
    g.manual_seed(1000 + lo)
    Gr = 0.3 + 0.01 * torch.rand((hi - lo, k), generator=g, device=dev, dtype=tdt)
    Gs = 0.3 + 0.01 * torch.rand((hi - lo, k), generator=g, device=dev, dtype=tdt)
    kr = torch.ones((hi - lo, 1), device=dev, dtype=tdt)
    tr = torch.ones((nI, 1), device=dev, dtype=tdt)
    return [Gs, Gr, Ls, Lr, kr, tr]


def run_ours(a):
    import torch
    import torch.distributed as dist
    from hpfrec_b200.engine import Engine
    from hpfrec_b200.loops import CudaLoops
    from hpfrec_b200 import dist as hdist

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    rb = 4 if a.dtype == "f32" else 8
    tdt = torch.float32 if rb == 4 else torch.float64
    nU, nI, nnz, k = a.nusers, a.nitems, a.nnz, a.k
    svi = bool(a.users_per_batch or a.items_per_batch)
    if svi and world > 1:
        raise SystemExit("the SVI configuration (C4) is single-GPU by specification (SURVEY §8e)")
    options = {}
    for opt in a.option:
        name, val = opt.split("=")
        options[name] = float(val)

    # ---- multi-GPU: prove the sharded data plane before timing it -----------------------------------
    parity = None
    if world > 1 and not a.no_parity_check:
        # the same exchange, schedule and synchronisation the timed run below resolves to (per-rank share of the users)
        parity = hdist.sharded_parity_check(local, options=options, overlap=hdist.resolve_overlap(nU // world))

    # ---- synthetic inputs, identical on every rank; each rank keeps its user range ---------------
    u, i, y = synth_coo_torch(nU, nI, nnz, dev, seed=42, alpha=a.alpha)
    y = y.to(tdt)
    cuts = hdist.plan_user_shards(u, nU, world)
    lo, hi = cuts[rank], cuts[rank + 1]
    lu, li, ly = hdist.shard_triples(u, i, y, lo, hi)
    lu, li, ly = lu.to(torch.int32).contiguous(), li.to(torch.int32).contiguous(), ly.contiguous()
    del u, i, y
    torch.cuda.empty_cache()
    loops = CudaLoops(rb == 4, device=local)
    state = initial_state(a, loops, lo, hi, dev)
    nUl = hi - lo
    nnz_local = int(lu.shape[0])

    stream = torch.cuda.current_stream()
    eng = Engine(nUl, nI, k, rb, local)
    if svi:
        eng.set_option("panel_mb", 1e9)   # minibatches are assembled from plain CSR / CSC orderings
    for name, val in options.items():
        eng.set_option(name, val)
    eng.set_hyper(0.3, 0.3, 1.0, 0.3, 0.3, 1.0)
    runner = None
    if world > 1:
        runner = hdist.ShardedLoop(eng, mode=os.environ.get("HPF_MULTI", "auto"), graph=os.environ.get("HPF_GRAPH", "1") == "1")
        stream = runner.stream
    eng.load_state(*state)
    eng.load_coo(lu, li, ly)
    ld = eng.ld
    engine_config = eng.describe()
    if runner is not None:
        engine_config["exchange"] = runner.describe()

    svi_state = None
    if svi:
        svi_state = dict(rng=np.random.default_rng(123), users=np.arange(nU, dtype=np.int64),
                         items=np.arange(nI, dtype=np.int64), epoch=0)

    def steps(n):
        if svi:
            loops.svi_epochs(eng, svi_state, n, nU, nI, a.users_per_batch, a.items_per_batch)
        elif runner is not None:
            runner.run(n)
        else:
            eng.step_full(n)

    def sync_all():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def launches_so_far():
        return eng.launch_count + (runner.replayed_launches if runner is not None else 0)

    # ---- warm-up, then EXACTLY K timed steps ----------------------------------------------------------
    steps(max(a.warmup, 3))
    sync_all()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
        time.sleep(0.3)
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    sync_all()
    l0 = launches_so_far()
    t_wall0 = time.time()
    ev0.record(stream)
    steps(a.steps)
    ev1.record(stream)
    sync_all()
    t_wall1 = time.time()
    launches = launches_so_far() - l0
    ms = torch.tensor([ev0.elapsed_time(ev1)], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        tl = torch.tensor([launches], device=dev, dtype=torch.int64)
        dist.all_reduce(tl)
        launches = int(tl.item())
    ms_total = float(ms.item())
    clocks = sampler.stop(t_wall0, t_wall1) if rank == 0 else None
    ms_per_step = ms_total / a.steps
    value = 1000.0 / ms_per_step

    # ---- per-kernel split (separate profiled pass: events between kernels serialise iterations) ------
    phases = None
    if world == 1 and not svi:
        eng.set_option("timing", 1)
        eng.step_full(max(3, min(a.steps, 10)))
        torch.cuda.synchronize()
        pm, pn = eng.phase_ms()
        eng.set_option("timing", 0)
        if pn > 0:
            phases = [x / pn for x in pm]

    e2e = None
    if not a.no_e2e and not svi:
        # ---- end to end through the C ABI with HOST buffers (pinned): every call uploads state + triples,
        # builds the orderings, runs K iterations and downloads the eight result arrays.
        hu = lu.cpu().pin_memory().numpy()
        hi_ = li.cpu().pin_memory().numpy()
        hy = ly.cpu().pin_memory().numpy()
        hstate = [(t if isinstance(t, torch.Tensor) else torch.from_numpy(t)).cpu().pin_memory().numpy() for t in state]
        outs = dict(Gamma_shp=(nUl, k), Gamma_rte=(nUl, k), Lambda_shp=(nI, k), Lambda_rte=(nI, k),
                    k_rte=(nUl, 1), t_rte=(nI, 1), Theta=(nUl, k), Beta=(nI, k))
        out_pinned = {key: torch.empty(shape, dtype=tdt).pin_memory() for key, shape in outs.items()}
        out_state = {key: t.numpy() for key, t in out_pinned.items()}
        eng.close()
        if runner is not None:
            runner.close()
        del lu, li, ly, state
        torch.cuda.empty_cache()
        h2d = hu.nbytes + hi_.nbytes + hy.nbytes + sum(x.nbytes for x in hstate)
        d2h = sum(x.nbytes for x in out_state.values())

        def e2e_call(legs=None):
            t = [time.time()]

            def lap():
                if legs is not None:
                    torch.cuda.synchronize()
                    t.append(time.time())
            e = Engine(nUl, nI, k, rb, local)
            for name, val in options.items():
                e.set_option(name, val)
            e.set_hyper(0.3, 0.3, 1.0, 0.3, 0.3, 1.0)
            r = hdist.ShardedLoop(e, mode=os.environ.get("HPF_MULTI", "auto"), graph=False) if world > 1 else None
            lap()
            e.load_state(*hstate)
            lap()
            e.load_coo(hu, hi_, hy)
            lap()
            if r is not None:
                r.run(a.steps)
            else:
                e.step_full(a.steps)
            lap()
            e.export_state(**out_state)
            lap()
            e.close()
            if r is not None:
                r.close()
            if legs is not None:
                for name, j in (("create_ms", 0), ("load_state_ms", 1), ("load_coo_ms", 2), ("iterate_ms", 3), ("export_ms", 4)):
                    legs[name] = round(1e3 * (t[j + 1] - t[j]), 2)

        e2e_call()  # warm-up (allocator, page-locking paths)
        sync_all()
        t0 = time.time()
        e2e_call()
        sync_all()
        t_e2e = torch.tensor([time.time() - t0], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t_e2e, op=dist.ReduceOp.MAX)
        t_e2e = float(t_e2e.item())
        legs = {}
        e2e_call(legs)   # a third call with a synchronize between the legs: where the time goes
        e2e = {"value": a.steps / t_e2e, "unit": UNIT, "h2d_bytes_per_step": int(h2d * world / a.steps),
               "d2h_bytes_per_step": int(d2h * world / a.steps), "seconds_per_call": t_e2e,
               "iterations_per_call": a.steps, "breakdown_rank0": legs,
               "what": "hpf_create+load_state+load_coo(host pinned)+%d iterations+export_state(host) per call; "
                       "bytes are per call / iterations, summed over ranks; breakdown from a separate call with a "
                       "device synchronize after every leg" % a.steps}
    else:
        eng.close()
        if runner is not None:
            runner.close()

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- roofline: SURVEY §8(d) algorithmic bytes of ONE iteration on one GPU / device time per iteration --
    peak, peak_src = measured_peak()
    b_iter = algorithmic_bytes(nUl, nI, nnz_local, k, rb)      # per GPU (item side replicated)
    achieved = b_iter / (ms_per_step * 1e-3) / 1e9
    wl_key = "%dx%dx%d k=%d %s" % (nU, nI, nnz, k, a.dtype)
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": None, "peak_source": peak_src,
                "kernel": "whole iteration: 2 x sweep_rows_kernel + 2 x update_rows_kernel" if not svi else
                          "whole SVI epoch (all minibatches)",
                "algorithmic_bytes_per_iteration_per_gpu": int(b_iter), "bytes_per_nnz": b_iter / max(1, nnz_local),
                "what": "SURVEY §8(d): nnz*(4+4+s) + 4*(nU+nI)*k*s bytes per iteration / device ms per iteration"}
    if phases is not None:
        tot = sum(phases)
        names = ("sweep_rows_kernel (item-major pass)", "sweep_rows_kernel (user-major pass)",
                 "update_rows_kernel (users)", "update_rows_kernel (items)")
        roofline["kernels"] = [{"name": nm, "ms": phases[j], "share_of_step": phases[j] / tot} for j, nm in enumerate(names)]
        roofline["dominant_kernel"] = {"name": "sweep_rows_kernel", "launches_per_iteration": 2,
                                       "ms_per_launch": (phases[0] + phases[1]) / 2,
                                       "share_of_step": (phases[0] + phases[1]) / tot}
        roofline["kernels_note"] = ("per-kernel times come from a separate pass with events between the kernels, which runs "
                                    "them one after the other; in the timed region the user update runs UNDER the item-major "
                                    "pass on a second stream (engine option overlap_update), so ms_per_step is %.3f ms less "
                                    "than their sum" % (tot - ms_per_step)) if tot > ms_per_step * 1.01 else None
        try:  # measured DRAM bytes per iteration from the committed ncu capture (same workload + configuration only)
            tr = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
            if tr.get("workload") == wl_key:
                roofline["traffic"] = int(tr["dram_bytes_per_sweep_launch"])
                roofline["traffic_per_iteration"] = int(tr["dram_bytes_per_iteration"])
                roofline["traffic_over_algorithmic"] = tr["dram_bytes_per_iteration"] / b_iter
                roofline["traffic_source"] = tr.get("source")
        except Exception:
            pass
        try:  # the measured ceiling of the bare row-gather pattern (tools/gather_probe.cu), for context
            gc = json.load(open(os.path.join(ROOT, "profiles", "gather_ceiling.json")))
            if gc.get("workload") == wl_key:
                roofline["gather_ceiling"] = {"ms_per_pass": gc["ms_per_pass"],
                                              "frac_of_ceiling": gc["ms_per_pass"] / roofline["dominant_kernel"]["ms_per_launch"],
                                              "source": gc["source"]}
        except Exception:
            pass

    metric, unit = (METRIC, UNIT) if not svi else ("svi_epochs_per_s", "epochs/s")
    line = {"metric": metric, "value": value, "unit": unit, "n_gpus": world, "steps": a.steps,
            "warmup": max(a.warmup, 3), "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": a.dtype, "data": "synthetic",
            "config": {"workload": workload_name(a), "parallelism": "user-sharded x%d, item side replicated%s" % (
                           world, "" if world == 1 else ", exchange=" + runner.mode),
                       "l2": "inputs larger than L2 (triples %.2f GB + factors %.2f GB per GPU)" % (
                           2 * (8 + rb) * nnz_local / 1e9, 4 * (nUl + nI) * ld * rb / 1e9),
                       "timing": "CUDA events on the engine stream, barrier+synchronize both sides, max over ranks",
                       "init": a.init,
                       "materialize": "shape/rate matrices stored on the last iteration of each call; "
                                      "intermediate iterations keep the equivalent per-row factors"},
            "nnz_per_s": value * nnz, "gpu_launches": int(launches), "clocks": clocks, "roofline": roofline,
            "engine": engine_config}
    if parity is not None:
        line["parity_check"] = parity
    if e2e is not None:
        line["e2e"] = e2e
    if not a.no_cpu_baseline and not svi:
        try:
            cb = cpu_reference_rate(a, 1, 3, max(1, a.cpu_sample_div))
        except Exception as exc:  # never lose the GPU line to a CPU-side failure
            cb = None
            line["cpu_baseline_error"] = repr(exc)
        if cb is not None:
            line["cpu_baseline"] = {"value": cb["nnz_per_s"] / a.nnz, "unit": UNIT, "cores": cb["cores"], "kind": "reference",
                                    "sample": cb["sample"], "nnz_per_s": cb["nnz_per_s"], "cpu_model": cb["cpu_model"],
                                    "note": "nnz/s of the sample scaled to the workload's nnz (the CPU path is compute-bound "
                                            "per (nnz, factor)); `bench.py --impl reference` times the full configuration"}
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    args = parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)
