#!/usr/bin/env python3
"""Build the UNMODIFIED reference hot path into oracle/_ref/  (TEST INFRASTRUCTURE, not product).

The reference (david-cortes/hpfrec) keeps its whole hot path -- the L1 C loops and the L2
iteration drivers `fit_hpf`, `partial_fit`, `initialize_parameters`, `calc_llk`, `predict_arr`,
`calc_user_factors` -- in ONE Cython text file, `hpfrec/cython_loops.pxi`, instantiated twice through
`hpfrec/cython_double_nonwindows.pyx` and `hpfrec/cython_float_nonwindows.pyx`
(reference setup.py:222-241).  This recipe

  1. runs `cython` on those two .pyx files *where they lie* under /root/reference (nothing is copied
     into this repository; the generated C goes to oracle/_ref/build/),
  2. compiles the generated C with /usr/bin/gcc and the flag set the reference's own setup.py ends up
     with on Linux (setup.py:39-46: -O2 -fopenmp -fno-math-errno -fno-trapping-math -std=c99), except
     that `-march=native` is replaced by `-march=x86-64-v3` because the .so is built in the CPU-only
     build container and must also run on the GPU box's (possibly different) host CPU,
  3. writes `oracle/_ref/hpfrec/cython_loops_{double,float}.*.so` (a namespace package: no __init__).

It does NOT run the reference's setup.py.  Trap recorded in SURVEY.md §8c: the image exports
CC=/opt/gcc/bin/gcc, which has no libgomp spec; with it -fopenmp fails and one silently gets a
single-threaded build, so /usr/bin/gcc is forced here and OpenMP is asserted after the build.

oracle/_ref/ is git-ignored (binaries stay out of history) but NOT gpurun-ignored, so the built
.so files travel to the GPU box, where /root/reference does not exist.

Usage:  python oracle/build_ref.py [--reference /root/reference] [--force]
"""
import argparse
import os
import subprocess
import sys
import sysconfig

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "_ref")
PKG = os.path.join(OUT, "hpfrec")
BUILD = os.path.join(OUT, "build")

MODULES = {
    # module name (as in reference setup.py:222-241)  ->  source file under <reference>/hpfrec/
    "hpfrec.cython_loops_double": "cython_double_nonwindows.pyx",
    "hpfrec.cython_loops_float": "cython_float_nonwindows.pyx",
}

CFLAGS = ["-O2", "-march=x86-64-v3", "-fopenmp", "-fno-math-errno", "-fno-trapping-math",
          "-std=c99", "-fPIC", "-w", "-DNPY_NO_DEPRECATED_API=NPY_1_7_API_VERSION"]


def so_path(modname):
    ext = sysconfig.get_config_var("EXT_SUFFIX")
    return os.path.join(PKG, modname.split(".")[-1] + ext)


def is_built():
    return all(os.path.exists(so_path(m)) for m in MODULES)


def build(reference="/root/reference", force=False, verbose=True):
    """Returns True if oracle/_ref is usable afterwards (built now or earlier)."""
    if is_built() and not force:
        return True
    src_dir = os.path.join(reference, "hpfrec")
    if not os.path.isdir(src_dir):
        if verbose:
            print("[oracle/build_ref] %s not present; keeping whatever is prebuilt" % src_dir)
        return is_built()
    import numpy as np
    os.makedirs(PKG, exist_ok=True)
    os.makedirs(BUILD, exist_ok=True)
    gcc = "/usr/bin/gcc" if os.path.exists("/usr/bin/gcc") else "gcc"
    pyinc = sysconfig.get_paths()["include"]
    for modname, pyx in MODULES.items():
        c_file = os.path.join(BUILD, modname.split(".")[-1] + ".c")
        cmd = [sys.executable, "-m", "cython", "-3", "--module-name", modname,
               "-o", c_file, os.path.join(src_dir, pyx)]
        if verbose:
            print("[oracle/build_ref]", " ".join(cmd))
        subprocess.check_call(cmd)
        cmd = [gcc, "-shared"] + CFLAGS + ["-I", pyinc, "-I", np.get_include(),
                                           c_file, "-o", so_path(modname), "-lm", "-fopenmp"]
        if verbose:
            print("[oracle/build_ref]", " ".join(cmd))
        subprocess.check_call(cmd)
    # the generated C is large (≈1.5 MB each) and not needed once compiled
    for f in os.listdir(BUILD):
        os.remove(os.path.join(BUILD, f))
    os.rmdir(BUILD)
    with open(os.path.join(OUT, "BUILD_INFO.txt"), "w") as fh:
        import scipy, Cython
        fh.write("reference: david-cortes/hpfrec @ %s\n" % reference)
        fh.write("numpy %s scipy %s cython %s python %s\n" % (
            np.__version__, scipy.__version__, Cython.__version__, sys.version.split()[0]))
        fh.write("cc: %s %s\n" % (gcc, " ".join(CFLAGS)))
    return True


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--reference", default="/root/reference")
    ap.add_argument("--force", action="store_true")
    a = ap.parse_args()
    ok = build(a.reference, a.force)
    print("[oracle/build_ref] built" if ok else "[oracle/build_ref] NOT built")
    sys.exit(0 if ok else 1)
