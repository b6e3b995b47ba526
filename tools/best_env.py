#!/usr/bin/env python3
"""Prints `export HPF_ROW_ALIGN=... HPF_OPTIONS=...` for one entry of gpurun_out/best.json (written by
tools/tune_r2.py), so that a shell can run the unmodified tests / bench under that configuration:

    eval "$(python tools/best_env.py H_k50_alpha0.6)"
"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def main():
    key = sys.argv[1] if len(sys.argv) > 1 else "H_k50_alpha0.6"
    path = os.path.join(ROOT, "gpurun_out", "best.json")
    try:
        entry = json.load(open(path))[key]
    except Exception as exc:  # no tuning result: leave the environment alone
        print("echo 'best_env: %s'" % str(exc).replace("'", ""))
        return
    print("export HPF_ROW_ALIGN=%s HPF_OPTIONS='%s'" % (entry["HPF_ROW_ALIGN"], entry["HPF_OPTIONS"]))


if __name__ == "__main__":
    main()
