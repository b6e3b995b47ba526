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
LIB = os.path.join(OUT_DIR, "libhpf_b200_tune.so" if os.environ.get("HPF_TUNE") else "libhpf_b200.so")
SOURCES = ["hpf_engine.cu"]
DEPS = ["hpf_engine.cu", "hpf_kernels.cuh", "hpf_batch.cuh", "hpf_device.cuh", "hpf_batch_host.inl",
        os.path.join("..", "..", "include", "hpf_b200.h")]

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17", "--expt-relaxed-constexpr", "--extended-lambda",
    "-Xcompiler", "-fPIC", "-Xcompiler", "-O2", "-shared",
    "-Xptxas", "-v" if os.environ.get("HPF_PTXAS_V") else "-O3",
] + (["-DHPF_TUNE"] if os.environ.get("HPF_TUNE") else [])


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


def build(force=False, verbose=False):
    if not force and not needs_build():
        return LIB
    os.makedirs(OUT_DIR, exist_ok=True)
    cmd = [find_nvcc()] + NVCC_FLAGS + [os.path.join(CSRC, s) for s in SOURCES] + ["-o", LIB]
    if verbose:
        print(" ".join(cmd), flush=True)
    # the image exports CC=/opt/gcc/bin/gcc (a wrapper); nvcc must use the system host compiler
    env = dict(os.environ)
    env.pop("CC", None)
    env.pop("CXX", None)
    res = subprocess.run(cmd + ["-ccbin", "/usr/bin/g++"] if os.path.exists("/usr/bin/g++") else cmd,
                         env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if verbose or res.returncode != 0:
        print(res.stdout)
    if res.returncode != 0:
        raise RuntimeError("nvcc failed building libhpf_b200.so")
    return LIB


if __name__ == "__main__":
    build(force="--force" in sys.argv, verbose=True)
    print("built", LIB)
