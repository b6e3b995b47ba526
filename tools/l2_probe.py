#!/usr/bin/env python3
"""DRAM traffic / L2 hit rate of the sweep kernel vs panel size and L2 hint (GPU box)."""
import csv, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
os.chdir(ROOT)
env = dict(os.environ)
METRICS = "dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct,gpu__time_duration.sum"
panels = [float(x) for x in (sys.argv[1].split(",") if len(sys.argv) > 1 else "12,24,32,48,96".split(","))]
for P in panels:
    for H in (0, 1):
        log = "gpurun_out/_l2probe.csv"
        cmd = ["ncu", "--metrics", METRICS, "--clock-control", "none", "-k", "regex:sweep_major", "-s", "4", "-c", "2",
               "--csv", "--log-file", log, sys.executable, "bench.py", "--steps", "2", "--warmup", "1", "--no-e2e",
               "--no-cpu-baseline", "--option", "panel_mb=%g" % P, "--option", "hint=%d" % H, "--option", "lpg=4",
               "--option", "unroll=1", "--option", "minb=3"]
        subprocess.run(cmd, env=env, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
        rows = [r for r in csv.reader(open(log)) if len(r) > 5]
        hdr = rows[0]
        ki, mi, vi, ui, ii = (hdr.index(x) for x in ("Kernel Name", "Metric Name", "Metric Value", "Metric Unit", "ID"))
        per = {}
        for r in rows[1:]:
            per.setdefault(r[ii], {})[r[mi]] = (r[vi], r[ui])
        for kid, m in per.items():
            print("panel_mb=%g hint=%d launch=%s " % (P, H, kid) + " ".join("%s=%s%s" % (k.split(".")[0].replace("__", "_"), v[0], v[1]) for k, v in sorted(m.items())), flush=True)
