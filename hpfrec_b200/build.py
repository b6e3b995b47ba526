"""Builds the engine's C-ABI shared library IN-TREE for sm_100a:

    hpfrec_b200/_lib/libhpf_b200.so   <-  hpfrec_b200/csrc/hpf_engine.cu (+ .cuh/.inl)

nvcc cross-compiles without a GPU.  The .so is git-ignored but travels to the GPU box with the
repository snapshot.  Usage: `python -m hpfrec_b200.build [--force] [--verbose]`.
"""
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OUT_DIR = os.path.join(HERE, "_lib")
LIB = os.path.join(OUT_DIR, "libhpf_b200.so")
PROBE_SRC = os.path.join(HERE, "..", "tools", "gather_probe.cu")
PROBE_BIN = os.path.join(HERE, "..", "tools", "bin", "gather_probe")
SOURCES = ["hpf_engine.cu"]
DEPS = ["hpf_engine.cu", "hpf_kernels.cuh", "hpf_batch.cuh", "hpf_device.cuh", "hpf_batch_host.inl", "hpf_sweep_dispatch.inl", "hpf_sweep.cuh", "hpf_scorer.inl", "hpf_ingest.inl",
        os.path.join("..", "..", "include", "hpf_b200.h")]

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17", "--expt-relaxed-constexpr", "--extended-lambda",
    "-Xcompiler", "-fPIC", "-Xcompiler", "-O2", "-shared",
    "-Xptxas", "-v" if os.environ.get("HPF_PTXAS_V") else "-O3",
]


def find_nvcc():
    for cand in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", shutil.which("nvcc")):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found: the HPF engine has no non-CUDA build")


def needs_build():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    return any(os.path.getmtime(os.path.join(CSRC, d)) > t for d in DEPS)


def _run_nvcc(cmd, verbose, what):
    # the image exports CC=/opt/gcc/bin/gcc (a wrapper); nvcc must use the system host compiler
    env = dict(os.environ)
    env.pop("CC", None)
    env.pop("CXX", None)
    if os.path.exists("/usr/bin/g++"):
        cmd = cmd + ["-ccbin", "/usr/bin/g++"]
    if verbose:
        print(" ".join(cmd), flush=True)
    res = subprocess.run(cmd, env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if verbose or res.returncode != 0:
        print(res.stdout)
    if res.returncode != 0:
        raise RuntimeError("nvcc failed building %s" % what)


def build_probe(force=False, verbose=False):
    """tools/bin/gather_probe: the measured ceiling of the sweep's row-gather pattern (a measurement
    tool, not part of the library)."""
    src, out = os.path.normpath(PROBE_SRC), os.path.normpath(PROBE_BIN)
    if not os.path.exists(src):
        return None
    if not force and os.path.exists(out) and os.path.getmtime(out) >= os.path.getmtime(src):
        return out
    os.makedirs(os.path.dirname(out), exist_ok=True)
    _run_nvcc([find_nvcc(), "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17", src,
               "-o", out], verbose, "tools/bin/gather_probe")
    return out


def build(force=False, verbose=False):
    if force or needs_build():
        os.makedirs(OUT_DIR, exist_ok=True)
        cmd = [find_nvcc()] + NVCC_FLAGS + [os.path.join(CSRC, s) for s in SOURCES] + ["-o", LIB]
        _run_nvcc(cmd, verbose, "libhpf_b200.so")
    build_probe(force=force, verbose=verbose)
    return LIB


if __name__ == "__main__":
    build(force="--force" in sys.argv, verbose=True)
    print("built", LIB)
