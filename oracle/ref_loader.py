"""Loader for the compiled UNMODIFIED reference hot path (oracle/_ref) -- TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / `--impl reference` legs may import
this.  The product package (hpfrec_b200) never does.

`load(use_float)` returns the compiled module `hpfrec.cython_loops_{float,double}` built by
oracle/build_ref.py from /root/reference/hpfrec/cython_{float,double}_nonwindows.pyx; its module-level
callables (`fit_hpf`, `partial_fit`, `initialize_parameters`, `calc_llk`, `predict_arr`,
`calc_user_factors`; reference hpfrec/cython_loops.pxi:117,147,423,476,525,538) are the reference's
own L2 drivers + L1 loops.  Returns None if oracle/_ref has not been built.
"""
import importlib
import os
import sys

_HERE = os.path.dirname(os.path.abspath(__file__))
_REF = os.path.join(_HERE, "_ref")


def available():
    d = os.path.join(_REF, "hpfrec")
    return os.path.isdir(d) and any(f.startswith("cython_loops_double") for f in os.listdir(d))


def load(use_float=False):
    if not available():
        return None
    if _REF not in sys.path:
        sys.path.insert(0, _REF)
    name = "hpfrec.cython_loops_float" if use_float else "hpfrec.cython_loops_double"
    return importlib.import_module(name)


def ref_fit_hpf(mod, Y, ix_u, ix_i, nU, nI, k, maxiter, seed=123, ncores=1, par_sh=0,
                users_per_batch=0, items_per_batch=0, step_size=None, st_ix_u=None,
                a=0.3, a_prime=0.3, b_prime=1.0, c=0.3, c_prime=0.3, d_prime=1.0,
                sum_exp_trick=0, alloc_full_phi=0):
    """Thin positional-argument adapter around the reference's `fit_hpf` (pxi:147-162) with
    stop_crit='maxiter', check_every=0, verbose=0, no validation set, keep_all_objs=1 -- the call
    `HPF._fit` makes (reference hpfrec/__init__.py:650-669).
    Returns dict(Theta, Beta, Gamma_shp, Gamma_rte, Lambda_shp, Lambda_rte, k_rte, t_rte, niter)."""
    import numpy as np
    real = mod.c_real_t
    ind = mod.obj_ind_type
    Theta = np.empty((nU, k), dtype=real)
    Beta = np.empty((nI, k), dtype=real)
    if step_size is None:
        step_size = lambda x: 1 / np.sqrt(x + 2)
    if st_ix_u is None:
        st_ix_u = np.arange(1).astype(ind)
    empty_r = np.empty(0, dtype=real)
    empty_i = np.empty(0, dtype=ind)
    niter, temp, llk = mod.fit_hpf(
        a, a_prime, b_prime, c, c_prime, d_prime,
        np.ascontiguousarray(Y, dtype=real), np.ascontiguousarray(ix_u, dtype=ind),
        np.ascontiguousarray(ix_i, dtype=ind), Theta, Beta,
        int(maxiter), "maxiter", 0, 1e-3, int(users_per_batch), int(items_per_batch),
        step_size, int(sum_exp_trick), np.ascontiguousarray(st_ix_u, dtype=ind),
        "", int(seed), 0, int(ncores), int(par_sh), 0, empty_r, empty_i, empty_i,
        0, 1, int(alloc_full_phi))
    return dict(Theta=Theta, Beta=Beta, Gamma_shp=temp[0], Gamma_rte=temp[1], Lambda_shp=temp[2],
                Lambda_rte=temp[3], k_rte=temp[4], t_rte=temp[5], niter=niter)
