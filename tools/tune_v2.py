#!/usr/bin/env python3
"""Tuning harness for the pipelined two-pass sweep kernel (option kernel=2) on the H workload (GPU box).

    python tools/tune_v2.py

Compares, on the same data and from the same state, with every result first verified against the
classic two-pass kernel (max relative difference of Theta / Beta after 2 iterations):
  * the classic two-pass kernel and the fused item-major pass (the leaders of tools/tune_r2.py),
  * the pipelined kernel over lane-group width, resident CTAs/SM, hints, chunk, L2 panel, row alignment,
and then the leaders on the hotter Zipf(0.9) data and at k=30 / k=128.
Output: gpurun_out/tune_v2.jsonl, gpurun_out/best.json (entries consumed by tools/best_env.py).
"""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import torch  # noqa: E402

import bench  # noqa: E402
from hpfrec_b200.engine import Engine  # noqa: E402
from hpfrec_b200.loops import CudaLoops  # noqa: E402

T0 = time.time()
NU, NI, NNZ = 1_000_000, 380_000, 48_000_000
if "--tiny" in sys.argv:
    NU, NI, NNZ = 20_000, 8_000, 400_000
dev = torch.device("cuda", 0)
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
out = open(os.path.join(ROOT, "gpurun_out", "tune_v2.jsonl"), "a")
loops = CudaLoops(True, device=0)
_data, _state = {}, {}


def get_data(alpha):
    if alpha not in _data:
        _data.clear()
        u, i, y = bench.synth_coo_torch(NU, NI, NNZ, dev, alpha=alpha)
        _data[alpha] = (u.to(torch.int32).contiguous(), i.to(torch.int32).contiguous(), y.contiguous())
    return _data[alpha]


def get_state(k):
    if k not in _state:
        _state.clear()
        st = loops.initialize_parameters(np.empty((NU, k), np.float32), np.empty((NI, k), np.float32),
                                         123, 0.3, 0.3, 1.0, 0.3, 0.3, 1.0)
        _state[k] = [torch.from_numpy(np.ascontiguousarray(x)).to(dev) for x in st]
    return _state[k]


class Setup:
    def __init__(self, k, alpha, align, panel_mb):
        self.k, self.alpha, self.align, self.panel_mb = k, alpha, align, panel_mb
        os.environ["HPF_ROW_ALIGN"] = str(align)
        os.environ.pop("HPF_OPTIONS", None)
        self.eng = Engine(NU, NI, k, 4, 0)
        self.eng.set_option("panel_mb", panel_mb)
        self.eng.set_option("strict", 1)
        u, i, y = get_data(alpha)
        self.eng.load_state(*get_state(k))
        self.eng.load_coo(u, i, y)
        self.theta = torch.empty((NU, k), dtype=torch.float32, device=dev)
        self.beta = torch.empty((NI, k), dtype=torch.float32, device=dev)

    def close(self):
        self.eng.close()

    def run(self, ref, tag, **opts):
        eng = self.eng
        rec = dict(tag=tag, k=self.k, alpha=self.alpha, row_align=self.align, ld=eng.ld, panel_mb=self.panel_mb, **opts)
        try:
            for name, val in opts.items():
                eng.set_option(name, val)
            eng.load_state(*get_state(self.k))
            eng.step_full(2)
            eng.export_state(Theta=self.theta, Beta=self.beta)
            if ref is not None:
                dth = float(((self.theta - ref[0]).abs() / ref[0].abs().clamp_min(1e-30)).max())
                dbe = float(((self.beta - ref[1]).abs() / ref[1].abs().clamp_min(1e-30)).max())
                rec["max_rel_diff"] = max(dth, dbe)
                rec["ok"] = bool(rec["max_rel_diff"] < 2e-4 and torch.isfinite(self.theta).all())
            eng.set_option("timing", 1)
            eng.step_full(3)
            torch.cuda.synchronize()
            ms, n = eng.phase_ms()
            eng.set_option("timing", 0)
            rec["ms"] = [round(x / n, 4) for x in ms]
            rec["ms_sweep"] = round((ms[0] + ms[1]) / n, 4)
            rec["ms_iter"] = round(sum(ms) / n, 4)
        except Exception as exc:
            rec["error"] = repr(exc)[:200]
            rec["ok"] = False
        rec["t"] = round(time.time() - T0, 1)
        out.write(json.dumps(rec) + "\n")
        out.flush()
        if ref is None:
            rec["ok"] = True
            return rec, (self.theta.clone(), self.beta.clone())
        return rec


def as_env(r):
    keys = ("sweep", "kernel", "lpg", "minb", "hint", "block", "chunk")
    opts = ["panel_mb=%g" % r["panel_mb"]] + ["%s=%d" % (kk, r[kk]) for kk in keys if kk in r]
    return {"HPF_ROW_ALIGN": str(r["row_align"]), "HPF_OPTIONS": ",".join(opts), "ms_iter": r.get("ms_iter"), "record": r}


def top(rows, n=10, f=lambda r: True):
    good = sorted([r for r in rows if r.get("ok") and "ms_iter" in r and f(r)], key=lambda r: r["ms_iter"])
    return good[:n]


CLASSIC = dict(sweep=0, kernel=1, lpg=4, minb=3, hint=1, chunk=64)
FUSED = dict(sweep=4, kernel=1, lpg=8, minb=3, hint=0, chunk=64)
# (lpg, minb, hint, block) of the deep-pipeline kernel
V3_CORE = [(8, 2, 0, 256), (8, 3, 0, 256), (8, 4, 0, 128), (8, 6, 0, 128), (4, 2, 0, 128), (4, 3, 0, 128),
           (16, 4, 0, 256), (16, 6, 0, 256)]
V4_CORE = [(8, 2, 256), (8, 3, 256), (8, 4, 128), (4, 2, 128), (4, 3, 128), (16, 4, 256)]
V2_CORE = [(8, 3, 0), (8, 2, 0), (8, 4, 0), (8, 5, 0), (8, 6, 0), (8, 3, 1), (16, 3, 0), (16, 4, 0), (16, 5, 0),
           (16, 6, 0), (4, 2, 0), (4, 3, 0), (16, 4, 1)]


def main():
    results, best = [], {}
    s0 = Setup(50, 0.6, 32, 48.0)
    rec0, ref = s0.run(None, "classic", **CLASSIC)
    results.append(rec0)
    print("classic:", json.dumps(rec0), flush=True)
    plan = [(128, 96.0), (128, 64.0), (128, 128.0), (128, 48.0)]
    if "--skip-v2" not in sys.argv:
        plan += [(128, 24.0), (32, 48.0), (32, 32.0)]
    for align, panel in plan:
        try:
            st = s0 if (align, panel) == (32, 48.0) else Setup(50, 0.6, align, panel)
        except Exception as exc:
            print("setup failed", align, panel, repr(exc)[:200], flush=True)
            continue
        results.append(st.run(ref, "classic", **CLASSIC))
        results.append(st.run(ref, "fused4", **FUSED))
        if "--skip-v2" not in sys.argv:
            for lpg, minb, hint in V2_CORE:
                for chunk in (64, 256):
                    results.append(st.run(ref, "v2", sweep=0, kernel=2, lpg=lpg, minb=minb, hint=hint, chunk=chunk))
        for lpg, minb, hint, block in V3_CORE:
            for chunk in (128, 256, 512):
                results.append(st.run(ref, "v3", sweep=0, kernel=3, lpg=lpg, minb=minb, hint=hint, block=block, chunk=chunk))
        for lpg, minb, block in V4_CORE:
            for chunk in (128, 256, 512):
                results.append(st.run(ref, "v4", sweep=0, kernel=4, lpg=lpg, minb=minb, hint=0, block=block, chunk=chunk))
        if st is not s0:
            st.close()
    s0.close()
    print("== H top 15 ==")
    for r in top(results, 15):
        print(json.dumps(r), flush=True)
    # chunk / panel refinement around the best pipelined shape
    lead = top(results, 1, lambda r: r.get("kernel") in (2, 3, 4))
    if lead:
        b = lead[0]
        for panel in sorted({b["panel_mb"], 40.0, 56.0}):
            try:
                st = Setup(50, 0.6, b["row_align"], panel)
            except Exception as exc:
                print("setup failed", panel, repr(exc)[:200], flush=True)
                continue
            for chunk in (32, 64, 128, 512, 1024):
                extra = {"block": b["block"]} if "block" in b else {}
                results.append(st.run(ref, "refine", sweep=0, kernel=b["kernel"], lpg=b["lpg"], minb=b["minb"], hint=b["hint"], chunk=chunk, **extra))
            st.close()
    h_top = top(results, 12)
    print("== H top 12 after refinement ==")
    for r in h_top:
        print(json.dumps(r), flush=True)
    best["H_k50_alpha0.6"] = as_env(h_top[0])
    for ver in (2, 3, 4):
        lead_v = top(results, 1, lambda r: r.get("kernel") == ver)
        if lead_v:
            best["H_k50_alpha0.6_v%d" % ver] = as_env(lead_v[0])
    best["H_k50_alpha0.6_classic"] = as_env(rec0)
    json.dump(best, open(os.path.join(ROOT, "gpurun_out", "best.json"), "w"), indent=1)

    # hotter item distribution
    leaders = []
    for r in top(results, 50):
        key = (r["sweep"], r.get("kernel", 1))
        if key not in [(x["sweep"], x.get("kernel", 1)) for x in leaders]:
            leaders.append(r)
    s9 = Setup(50, 0.9, 32, 48.0)
    rec9, ref9 = s9.run(None, "classic-alpha0.9", **CLASSIC)
    s9.close()
    res9 = [rec9]
    for r in leaders[:3]:
        st = Setup(50, 0.9, r["row_align"], r["panel_mb"])
        res9.append(st.run(ref9, "leader-alpha0.9", **{kk: r[kk] for kk in ("sweep", "kernel", "lpg", "minb", "hint", "block", "chunk") if kk in r}))
        st.close()
    print("== alpha 0.9 ==")
    for r in res9:
        print(json.dumps(r), flush=True)
    best["H_k50_alpha0.9_all"] = [as_env(r) for r in top(res9, 5)]
    json.dump(best, open(os.path.join(ROOT, "gpurun_out", "best.json"), "w"), indent=1)

    # other row classes: classic default vs fused vs pipelined candidates
    for k, classic, fused, v2shapes in (
            (30, dict(lpg=4, minb=2, hint=0), dict(lpg=8, minb=6, hint=0), [(8, 4, 0), (8, 6, 0), (8, 8, 0), (4, 3, 0), (4, 4, 0), (4, 6, 0)]),
            (128, dict(lpg=8, minb=4, hint=0), dict(lpg=16, minb=4, hint=0), [(16, 3, 0), (16, 2, 0), (16, 4, 0), (32, 3, 0), (32, 4, 0), (8, 2, 0), (8, 3, 0)])):
        resk, refk = [], None
        for align, panel in ((128, 48.0), (128, 96.0)):
            st = Setup(k, 0.6, align, panel)
            if refk is None:
                r, refk = st.run(None, "classic-k%d" % k, sweep=0, kernel=1, chunk=64, **classic)
                resk.append(r)
            else:
                resk.append(st.run(refk, "classic-k%d" % k, sweep=0, kernel=1, chunk=64, **classic))
            resk.append(st.run(refk, "fused4-k%d" % k, sweep=4, kernel=1, chunk=64, **fused))
            if "--skip-v2" not in sys.argv:
                for lpg, minb, hint in v2shapes:
                    resk.append(st.run(refk, "v2-k%d" % k, sweep=0, kernel=2, lpg=lpg, minb=minb, hint=hint, chunk=64))
            for lpg, minb, block in ([(4, 2, 256), (4, 3, 256), (4, 6, 128), (8, 4, 256), (8, 6, 256)] if k == 30 else
                                     [(8, 2, 128), (8, 3, 128), (16, 2, 256), (16, 3, 256), (32, 4, 256), (32, 6, 256)]):
                for chunk in (64, 256):
                    resk.append(st.run(refk, "v3-k%d" % k, sweep=0, kernel=3, lpg=lpg, minb=minb, hint=0, block=block, chunk=chunk))
            for lpg, minb, block in ([(4, 2, 256), (4, 3, 256), (8, 4, 256)] if k == 30 else
                                     [(8, 3, 128), (8, 2, 128), (16, 2, 256), (16, 3, 256)]):
                resk.append(st.run(refk, "v4-k%d" % k, sweep=0, kernel=4, lpg=lpg, minb=minb, hint=0, block=block, chunk=256))
            st.close()
        print("== k=%d top 6 ==" % k)
        for r in top(resk, 6):
            print(json.dumps(r), flush=True)
        best["k%d_alpha0.6_all" % k] = [as_env(r) for r in top(resk, 6)]
        json.dump(best, open(os.path.join(ROOT, "gpurun_out", "best.json"), "w"), indent=1)
    out.close()
    print("tune_v2 done in %.0f s" % (time.time() - T0), flush=True)


if __name__ == "__main__":
    main()
