#!/usr/bin/env python3
"""DRAM traffic and L2 hit rate of the sweep / update kernels per (L2 panel size, hint mode, shape).

Run UNDER ncu on the GPU box (one process, data generated once):

    ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct,gpu__time_duration.sum \
        --clock-control none -k regex:'sweep_rows|update_rows' --csv --log-file gpurun_out/traffic.csv \
        python tools/traffic_probe.py

Every configuration runs 3 iterations (12 profiled launches: item-major pass, user-major pass, user update,
item update, three times; the middle one -- a lean iteration -- is reported); the order of configurations is written to gpurun_out/traffic_configs.json so that
tools/traffic_probe.py --summarise can join the ncu CSV with it (works offline).
"""
import csv
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
OUT = os.path.join(ROOT, "gpurun_out")


def configs():
    """The production configuration at three L2 panel sizes, and the register form beside it."""
    out = []
    for panel in (96.0, 64.0, 48.0, 128.0):
        out.append(dict(panel_mb=panel, smem_gather=1, fullrow=1, hint=0, chunk=256))
    out.append(dict(panel_mb=96.0, smem_gather=0, fullrow=0, hint=0, chunk=256, lpg=8, depth=2, block=256, minb=4))
    return out


def run():
    import numpy as np
    import torch
    import bench
    from hpfrec_b200.engine import Engine
    from hpfrec_b200.loops import CudaLoops
    nU, nI, nnz, k = 1_000_000, 380_000, 48_000_000, 50
    dev = torch.device("cuda", 0)
    u, i, y = bench.synth_coo_torch(nU, nI, nnz, dev, alpha=0.6)
    u, i, y = u.to(torch.int32).contiguous(), i.to(torch.int32).contiguous(), y.contiguous()
    loops = CudaLoops(True, device=0)
    st = loops.initialize_parameters(np.empty((nU, k), np.float32), np.empty((nI, k), np.float32), 123, 0.3, 0.3, 1.0, 0.3, 0.3, 1.0)
    st = [torch.from_numpy(np.ascontiguousarray(x)).to(dev) for x in st]
    done = []
    last_panel, eng = None, None
    for cfg in configs():
        if cfg["panel_mb"] != last_panel:
            if eng is not None:
                eng.close()
            eng = Engine(nU, nI, k, 4, 0)
            eng.set_option("panel_mb", cfg["panel_mb"])
            eng.load_state(*st)
            eng.load_coo(u, i, y)
            last_panel = cfg["panel_mb"]
        for name in ("lpg", "depth", "block", "minb"):
            eng.set_option(name, 0)
        for name, val in cfg.items():
            if name != "panel_mb":
                eng.set_option(name, val)
        eng.step_full(3)   # 2 lean iterations + the materialising one
        torch.cuda.synchronize()
        done.append(cfg)
        json.dump(done, open(os.path.join(OUT, "traffic_configs.json"), "w"))
    eng.close()


def summarise(csv_path, cfg_path):
    cfgs = json.load(open(cfg_path))
    rows = [r for r in csv.reader(open(csv_path)) if len(r) > 5]
    hdr = rows[0]
    ki, mi, vi, ii = (hdr.index(x) for x in ("Kernel Name", "Metric Name", "Metric Value", "ID"))
    per = {}
    for r in rows[1:]:
        d = per.setdefault(int(r[ii]), {"kernel": r[ki].split("(")[0].replace("void ", "")[:48]})
        d[r[mi]] = float(r[vi].replace(",", ""))
    ids = sorted(per)
    out = []
    for n, cfg in enumerate(cfgs):
        launches = [per[j] for j in ids[n * 12 + 4:n * 12 + 8]]   # second (lean) iteration of the configuration
        if len(launches) < 4:
            break
        rec = dict(cfg)
        names = ("item_major_pass", "user_major_pass", "update_users", "update_items")
        tot = 0.0
        for nm, l in zip(names, launches):
            b = l.get("dram__bytes_read.sum", 0.0) + l.get("dram__bytes_write.sum", 0.0)
            tot += b
            rec[nm] = {"dram_read": l.get("dram__bytes_read.sum"), "dram_write": l.get("dram__bytes_write.sum"),
                       "l2_hit_pct": l.get("lts__t_sector_hit_rate.pct"), "ncu_ns": l.get("gpu__time_duration.sum")}
        rec["dram_bytes_per_iteration"] = tot
        out.append(rec)
        print(json.dumps(rec))
    return out


if __name__ == "__main__":
    if "--summarise" in sys.argv:
        summarise(os.path.join(OUT, "traffic.csv"), os.path.join(OUT, "traffic_configs.json"))
    else:
        run()
