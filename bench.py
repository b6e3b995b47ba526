#!/usr/bin/env python3
"""bench.py -- full-batch CAVI iterations/s on the MillionSong-shaped synthetic (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W [--impl reference]

A "step" is one complete full-batch CAVI iteration (the reference's loop body, cython_loops.pxi:230-259)
over the whole nnz list.  Default workload = H of SURVEY.md §8: 1M x 380K users x items, 48M nnz,
k=50, fp32.  For N>1 the nnz list is sharded by user range (balanced by nnz) with the item side
replicated and one all-reduce per iteration; the problem size is fixed, so scaling is "strong".

Rank 0 prints ONE JSON line (see the contract in the task description): value = iterations/s from
device-resident inputs (CUDA events on the engine's stream, barrier + synchronize on both sides, max
over ranks); `e2e` = the same iterations through the C ABI with HOST (pinned) buffers: upload of state
and triples, index build, K iterations, download of the eight result arrays, all inside the timed
region; `roofline` against MEASURED_PEAKS.json; `cpu_baseline` = the compiled, unmodified reference
(oracle/_ref) timed on this box's host cores on a bounded sample.
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


# -----------------------------------------------------------------------------------------------------
def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--nusers", type=int, default=1_000_000)
    ap.add_argument("--nitems", type=int, default=380_000)
    ap.add_argument("--nnz", type=int, default=48_000_000)
    ap.add_argument("--k", type=int, default=50)
    ap.add_argument("--dtype", default="f32", choices=["f32", "f64"])
    ap.add_argument("--alpha", type=float, default=0.6, help="Zipf exponent of item popularity")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cpu-sample-div", type=int, default=24, help="CPU sample = workload / this")
    ap.add_argument("--option", action="append", default=[], help="engine option name=value")
    return ap.parse_args()


def workload_name(a):
    return "%dx%d users x items, %d nnz, k=%d, %s full-batch CAVI (synthetic: lognormal users, Zipf(%.1f) items)" % (
        a.nusers, a.nitems, a.nnz, a.k, "fp32" if a.dtype == "f32" else "fp64", a.alpha)


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


# -----------------------------------------------------------------------------------------------------
def cpu_reference_rate(a, steps, warmup, verbose=False):
    """Times the compiled UNMODIFIED reference (oracle/_ref: hpfrec.cython_loops_float.fit_hpf, all host
    threads, deterministic scatter) on a bounded sample of the workload.  The CPU path is compute-bound
    on psi/log/exp per (nnz, factor), so nnz/s is size-independent and the sample rate
    extrapolates to the full nnz list.  Returns dict or None if oracle/_ref is absent."""
    from oracle import ref_loader as R
    from oracle import hpf_oracle as O
    use_float = a.dtype == "f32"
    mod = R.load(use_float)
    if mod is None:
        return None
    div = max(1, a.cpu_sample_div)
    nU, nI, nnz = max(64, a.nusers // div), max(64, a.nitems // div), max(1024, a.nnz // div)
    u, i, y = O.synth_coo(nU, nI, nnz, seed=42, alpha=a.alpha)
    cores = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    dt = np.float32 if use_float else np.float64

    def run(iters):
        t0 = time.time()
        R.ref_fit_hpf(mod, y.astype(dt), u, i, nU, nI, a.k, iters, seed=123, ncores=cores, par_sh=0)
        return time.time() - t0

    t_base = run(1)                      # init + phi allocation + 1 iteration
    if warmup > 0:
        run(1)
    n_it = max(1, min(steps, 3))
    t_more = run(1 + n_it)
    s_per_iter = max(1e-9, (t_more - t_base) / n_it)
    nnz_per_s = nnz / s_per_iter
    model = ""
    try:
        for line in open("/proc/cpuinfo"):
            if line.startswith("model name"):
                model = line.split(":", 1)[1].strip()
                break
    except Exception:
        pass
    return {"nnz_per_s": nnz_per_s, "iters_per_s_full_workload": nnz_per_s / a.nnz, "cores": cores,
            "cpu_model": model, "s_per_iter_sample": s_per_iter,
            "sample": "%dx%d, %d nnz (workload/%d, same generator), k=%d, %s, (t[maxiter=%d]-t[maxiter=1])/%d"
                      % (nU, nI, nnz, div, a.k, "fp32" if use_float else "fp64", 1 + n_it, n_it)}


def run_reference(a):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    res = cpu_reference_rate(a, a.steps, a.warmup)
    if res is None:
        print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref not built on this box"}))
        return
    v = res["iters_per_s_full_workload"]
    line = {"impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": a.gpus,
            "steps": a.steps, "warmup": a.warmup, "ms_per_step": 1000.0 / v, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": a.dtype, "data": "synthetic",
            "config": {"workload": workload_name(a)}, "nnz_per_s": res["nnz_per_s"],
            "cpu_baseline": {"value": v, "unit": UNIT, "cores": res["cores"], "kind": "reference",
                             "sample": res["sample"], "cpu_model": res["cpu_model"]},
            "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line))


# -----------------------------------------------------------------------------------------------------
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
    npdt = np.float32 if rb == 4 else np.float64
    tdt = torch.float32 if rb == 4 else torch.float64
    nU, nI, nnz, k = a.nusers, a.nitems, a.nnz, a.k

    # ---- synthetic inputs, identical on every rank; each rank keeps its user range ---------------
    u, i, y = synth_coo_torch(nU, nI, nnz, dev, seed=42, alpha=a.alpha)
    y = y.to(tdt)
    cuts = hdist.plan_user_shards(u, nU, world)
    lo, hi = cuts[rank], cuts[rank + 1]
    lu, li, ly = hdist.shard_triples(u, i, y, lo, hi)
    lu, li, ly = lu.contiguous(), li.contiguous(), ly.contiguous()
    del u, i, y
    torch.cuda.empty_cache()
    loops = CudaLoops(rb == 4, device=local)
    Theta = np.empty((nU, k), npdt)
    Beta = np.empty((nI, k), npdt)
    Gs, Gr, Ls, Lr, kr, tr = loops.initialize_parameters(Theta, Beta, 123, 0.3, 0.3, 1.0, 0.3, 0.3, 1.0)
    del Theta, Beta
    Gs, Gr, kr = (np.ascontiguousarray(x[lo:hi]) for x in (Gs, Gr, kr))
    nUl = hi - lo

    stream = torch.cuda.current_stream()
    eng = Engine(nUl, nI, k, rb, local)
    for opt in a.option:
        name, val = opt.split("=")
        eng.set_option(name, float(val))
    eng.set_hyper(0.3, 0.3, 1.0, 0.3, 0.3, 1.0)
    eng.load_state(Gs, Gr, Ls, Lr, kr, tr)
    eng.load_coo(lu, li, ly)
    ld = eng.ld
    engine_config = eng.describe()

    if world > 1:
        partial = hdist.engine_partial_tensors(eng, local)

        mode = os.environ.get("HPF_MULTI", "peer")
        graphed = None
        if os.environ.get("HPF_GRAPH", "0") == "1":
            graphed = hdist.GraphedShardLoop(eng, mode)
            stream = graphed.stream
        elif mode == "peer":
            hdist.attach_peers(eng)

        def steps(n):
            if graphed is not None:
                with torch.cuda.stream(graphed.stream):
                    graphed.run(n)
            elif mode == "plain":
                hdist.run_sharded_iterations(eng, n, partial)
            elif mode == "overlap":
                hdist.run_sharded_iterations_overlapped(eng, n, partial)
            else:
                hdist.run_sharded_iterations_peer(eng, n)
    else:
        def steps(n):
            eng.step_full(n)

    if world == 1:
        graphed = None

    def sync_all():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- warm-up, then EXACTLY K timed steps ----------------------------------------------------------
    steps(max(a.warmup, 3))
    sync_all()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
        time.sleep(0.3)
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    sync_all()
    l0 = eng.launch_count + (graphed.replayed_launches if world > 1 and graphed is not None else 0)
    t_wall0 = time.time()
    ev0.record(stream)
    steps(a.steps)
    ev1.record(stream)
    sync_all()
    t_wall1 = time.time()
    launches = eng.launch_count + (graphed.replayed_launches if world > 1 and graphed is not None else 0) - l0
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
    if world == 1:
        eng.set_option("timing", 1)
        eng.step_full(max(3, min(a.steps, 10)))
        torch.cuda.synchronize()
        pm, pn = eng.phase_ms()
        eng.set_option("timing", 0)
        if pn > 0:
            phases = [x / pn for x in pm]

    out_state = None
    e2e = None
    if not a.no_e2e:
        # ---- end to end through the C ABI with HOST buffers (pinned): every call uploads state + triples,
        # builds the orderings, runs K iterations and downloads the eight result arrays.
        hu = lu.to(torch.int32).cpu().pin_memory().numpy()
        hi_ = li.to(torch.int32).cpu().pin_memory().numpy()
        hy = ly.cpu().pin_memory().numpy()
        pinned = [torch.from_numpy(x).pin_memory() for x in (Gs, Gr, Ls, Lr, kr, tr)]
        hstate = [t.numpy() for t in pinned]
        outs = dict(Gamma_shp=(nUl, k), Gamma_rte=(nUl, k), Lambda_shp=(nI, k), Lambda_rte=(nI, k),
                    k_rte=(nUl, 1), t_rte=(nI, 1), Theta=(nUl, k), Beta=(nI, k))
        out_pinned = {key: torch.empty(shape, dtype=tdt).pin_memory() for key, shape in outs.items()}
        out_state = {key: t.numpy() for key, t in out_pinned.items()}
        eng.close()
        del lu, li, ly
        torch.cuda.empty_cache()
        h2d = hu.nbytes + hi_.nbytes + hy.nbytes + sum(x.nbytes for x in hstate)
        d2h = sum(x.nbytes for x in out_state.values())

        def e2e_call():
            e = Engine(nUl, nI, k, rb, local)
            for opt in a.option:
                name, val = opt.split("=")
                e.set_option(name, float(val))
            e.set_hyper(0.3, 0.3, 1.0, 0.3, 0.3, 1.0)
            e.load_state(*hstate)
            e.load_coo(hu, hi_, hy)
            if world > 1 and os.environ.get("HPF_MULTI", "peer") == "peer":
                hdist.attach_peers(e)
                hdist.run_sharded_iterations_peer(e, a.steps)
            elif world > 1:
                hdist.run_sharded_iterations_overlapped(e, a.steps, hdist.engine_partial_tensors(e, local))
            else:
                e.step_full(a.steps)
            e.export_state(**out_state)
            n_l = e.launch_count
            e.close()
            return n_l

        e2e_call()  # warm-up (allocator, page-locking paths)
        sync_all()
        t0 = time.time()
        e2e_call()
        sync_all()
        t_e2e = torch.tensor([time.time() - t0], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t_e2e, op=dist.ReduceOp.MAX)
        t_e2e = float(t_e2e.item())
        e2e = {"value": a.steps / t_e2e, "unit": UNIT, "h2d_bytes_per_step": int(h2d * world / a.steps),
               "d2h_bytes_per_step": int(d2h * world / a.steps), "seconds_per_call": t_e2e,
               "iterations_per_call": a.steps,
               "what": "hpf_create+load_state+load_coo(host pinned)+%d iterations+export_state(host) per call; "
                       "bytes are per call / iterations, summed over ranks" % a.steps}
    else:
        eng.close()

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- roofline (whole iteration and dominant kernel) ---------------------------------------------------
    peak, peak_src = measured_peak()
    nnz_local_max = nnz / world  # shards are nnz-balanced
    b_iter = algorithmic_bytes(nUl, nI, nnz_local_max, k, rb)      # per GPU (item side replicated)
    achieved = b_iter / (ms_per_step * 1e-3) / 1e9
    iteration = {"achieved": achieved, "frac": achieved / peak,
                 "algorithmic_bytes_per_iteration_per_gpu": int(b_iter), "bytes_per_nnz": b_iter / nnz_local_max,
                 "what": "whole iteration (2 sweep passes + 2 row updates): nnz*(4+4+s)+4*(nU+nI)*k*s bytes / device ms"}
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": None, "peak_source": peak_src, "kernel": "whole iteration", "iteration": iteration}
    if phases is not None:
        s = rb
        sweep_kernel = engine_config.get("kernel", "sweep kernel")
        # per-launch algorithmic bytes of one pass: triples once + own x read + gathered x read once + sums written
        pass_bytes = [nnz * (4 + 4 + s) + (2 * nI + nUl) * k * s, nnz * (4 + 4 + s) + (2 * nUl + nI) * k * s]
        # a one-pass (fused) sweep streams the triples once and touches all four factor / sum matrices once
        fused_bytes = nnz * (4 + 4 + s) + 2 * (nUl + nI) * k * s
        upd_bytes = [4 * nUl * k * s, 4 * nI * k * s]  # x r/w, sums r/w(zero) (lean iteration)
        tot = sum(phases)
        live = [j for j in (0, 1) if phases[j] > 0.02]  # a fused sweep leaves the other pass's slot empty
        kernels = []
        for j in live:
            nbytes = pass_bytes[j] if len(live) == 2 else fused_bytes
            label = ("item-major pass", "user-major pass")[j] if len(live) == 2 else "one fused %s pass" % ("item-major", "user-major")[j]
            kernels.append({"name": "%s (%s)" % (sweep_kernel, label), "ms": phases[j], "share": phases[j] / tot,
                            "algorithmic_bytes": int(nbytes), "achieved_gbs": nbytes / (phases[j] * 1e-3) / 1e9,
                            "frac": nbytes / (phases[j] * 1e-3) / 1e9 / peak})
        for j, nm in ((2, "update_rows_kernel (users)"), (3, "update_rows_kernel (items)")):
            kernels.append({"name": nm, "ms": phases[j], "share": phases[j] / tot, "algorithmic_bytes": int(upd_bytes[j - 2]),
                            "achieved_gbs": upd_bytes[j - 2] / (phases[j] * 1e-3) / 1e9,
                            "frac": upd_bytes[j - 2] / (phases[j] * 1e-3) / 1e9 / peak})
        roofline["kernels"] = kernels
        # the contract's roofline object describes the DOMINANT kernel, per launch
        dom = kernels[:len(live)]
        dom_bytes = sum(kk["algorithmic_bytes"] for kk in dom) / len(dom)
        dom_ms = sum(kk["ms"] for kk in dom) / len(dom)
        dom_achieved = dom_bytes / (dom_ms * 1e-3) / 1e9
        roofline.update({"kernel": "%s (%d launch%s per iteration; per-launch averages)" % (
                             sweep_kernel, len(dom), "es" if len(dom) > 1 else ""),
                         "achieved": dom_achieved, "frac": dom_achieved / peak,
                         "algorithmic_bytes_per_launch": int(dom_bytes), "ms_per_launch": dom_ms,
                         "share_of_step": sum(kk["ms"] for kk in dom) / tot})
        try:  # measured DRAM bytes per launch of that kernel from the committed ncu capture (same workload + kernel only)
            tr = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
            if tr.get("workload") == "%dx%dx%d k=%d %s" % (nU, nI, nnz, k, a.dtype) and tr.get("kernel") == sweep_kernel:
                sw = tr["sweep_launches"]
                roofline["traffic"] = int(sum(v["dram_read_bytes"] + v["dram_write_bytes"] for v in sw) / len(sw))
                roofline["traffic_source"] = "profiles/traffic.json (ncu dram__bytes_read.sum + dram__bytes_write.sum per launch)"
        except Exception:
            pass
        try:  # the measured ceiling of the bare row-gather pattern (tools/gather_probe.cu), for context
            gc = json.load(open(os.path.join(ROOT, "profiles", "gather_ceiling.json")))
            if gc.get("workload") == "%dx%dx%d k=%d %s" % (nU, nI, nnz, k, a.dtype):
                roofline["gather_ceiling"] = {"ms_per_pass": gc["ms_per_pass"], "frac_of_ceiling": gc["ms_per_pass"] / dom_ms,
                                              "source": gc["source"]}
        except Exception:
            pass

    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": a.steps,
            "warmup": max(a.warmup, 3), "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": a.dtype, "data": "synthetic",
            "config": {"workload": workload_name(a), "parallelism": "user-sharded x%d, item side replicated%s" % (
                           world, "" if world == 1 else ", exchange=" + os.environ.get("HPF_MULTI", "peer")),
                       "l2": "inputs larger than L2 (triples %.2f GB + factors %.2f GB per GPU)" % (
                           2 * 12 * nnz_local_max / 1e9, 4 * (nUl + nI) * ld * rb / 1e9),
                       "timing": "CUDA events on the engine stream, barrier+synchronize both sides, max over ranks",
                       "materialize": "shape/rate matrices stored on the last iteration of each call; "
                                      "intermediate iterations keep the equivalent per-row factors" if world == 1
                                      else "every iteration"},
            "nnz_per_s": value * nnz, "gpu_launches": int(launches), "clocks": clocks, "roofline": roofline,
            "engine": engine_config}
    if e2e is not None:
        line["e2e"] = e2e
    if not a.no_cpu_baseline:
        try:
            cb = cpu_reference_rate(a, 3, 1)
        except Exception as exc:  # never lose the GPU line to a CPU-side failure
            cb = None
            line["cpu_baseline_error"] = repr(exc)
        if cb is not None:
            line["cpu_baseline"] = {"value": cb["iters_per_s_full_workload"], "unit": UNIT, "cores": cb["cores"],
                                    "kind": "reference", "sample": cb["sample"], "nnz_per_s": cb["nnz_per_s"],
                                    "cpu_model": cb["cpu_model"]}
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    args = parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)
