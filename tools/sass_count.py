#!/usr/bin/env python3
"""Offline instruction accounting of the sweep kernels (no GPU needed).

    python tools/sass_count.py [harness.cu]

Compiles hpfrec_b200/csrc/experimental/sass_harness.cu for sm_100a, disassembles every kernel in it and
prints, per kernel: registers, total SASS instructions, instructions per pipelined step (distance between
consecutive DEPBAR.LE = cp.async.wait_group) and the opcode mix of one steady-state step.  The step
length includes the two divergent blocks (own-row staging, major-id change), which the common path
skips; they are reported separately when recognisable (the REDG block).
"""
import collections
import os
import re
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = sys.argv[1] if len(sys.argv) > 1 and not sys.argv[1].startswith("--") else os.path.join(ROOT, "hpfrec_b200", "csrc", "experimental", "sass_harness.cu")


def main():
    with tempfile.TemporaryDirectory() as tmp:
        cubin = os.path.join(tmp, "h.cubin")
        cmd = ["/usr/local/cuda/bin/nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-std=c++17", "-lineinfo",
               "--expt-relaxed-constexpr", "-cubin", "-Xptxas", "-v", "-ccbin", "/usr/bin/g++", SRC, "-o", cubin]
        env = dict(os.environ)
        env.pop("CC", None)
        env.pop("CXX", None)
        res = subprocess.run(cmd, env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
        if res.returncode != 0:
            print(res.stdout)
            raise SystemExit("nvcc failed")
        regs = dict(re.findall(r"Function properties for (\S+)\n.*\n?ptxas info\s+: Used (\d+) registers", res.stdout))
        sass = subprocess.run(["/usr/local/cuda/bin/cuobjdump", "-sass", cubin], stdout=subprocess.PIPE, text=True).stdout
    kernels, name = {}, None
    for line in sass.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            name = m.group(1)
            kernels[name] = []
            continue
        m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(.*?);", line)
        if m and name:
            kernels[name].append(m.group(1).strip())
    for name, ins in kernels.items():
        short = subprocess.run(["/usr/local/cuda/bin/cu++filt", name], stdout=subprocess.PIPE, text=True).stdout.strip()
        short = short[:short.rfind(">(") + 1] if ">(" in short else short
        waits = [i for i, s in enumerate(ins) if s.startswith("DEPBAR.LE")]
        print("== %s" % short)
        print("   registers %s, SASS instructions %d, pipelined steps found %d" % (regs.get(name, "?"), len(ins), len(waits)))
        if len(waits) >= 4:
            a, b = waits[2], waits[3]
            step = ins[a:b]
            ops = collections.Counter(re.sub(r"^@!?U?P\w+\s+", "", s).split()[0].split(".")[0] for s in step)
            red = [i for i, s in enumerate(step) if "REDG" in s]
            print("   one steady-state step: %d instructions (incl. divergent blocks; REDG block spans ~%d)" % (
                len(step), (red[-1] - red[0] + 25) if red else 0))
            print("   mix: " + ", ".join("%s %d" % kv for kv in ops.most_common(18)))
            if "--dump" in sys.argv:
                for s_ in step:
                    print("      " + s_)


if __name__ == "__main__":
    main()
