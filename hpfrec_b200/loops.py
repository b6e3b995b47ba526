"""Host-side mirror of the reference's compiled modules `hpfrec.cython_loops_{float,double}`.

The reference's Python class talks to its hot path only through the module-level callables of those
two modules (call sites hpfrec/__init__.py:650-669, 882, 914-927, 1038, 1145, 1284,
1433).  `cuda_loops_float` / `cuda_loops_double` below export the same names with the same argument
lists, argument meaning and in-place/return conventions, but run the work on the B200 through
libhpf_b200.so.  Iteration *control* stays in Python exactly as it does in the reference (its L2 is
Python-level Cython); every numeric loop is a CUDA kernel.  There is no CPU path here.

Arguments that only steer the reference's CPU implementation are accepted and ignored, each provably
inside tolerance of the corresponding reference setting: `nthreads`, `par_sh`
(the GPU scatter never loses updates), `alloc_full_phi` (phi is never materialised),
`sum_exp_trick` (the engine always normalises per row before exponentiating).
"""
import ctypes
import os
import time

import numpy as np

from .engine import Engine, as_index


def _rows_in_order(indptr, ids):
    """Positions of all CSR entries of rows `ids`, row after row (vectorised get_i_batch_pass1/2,
    reference pxi:774-797)."""
    cnt = indptr[ids + 1] - indptr[ids]
    total = int(cnt.sum())
    if total == 0:
        return np.empty(0, dtype=np.int64), cnt
    starts = np.repeat(indptr[ids] - np.concatenate(([0], np.cumsum(cnt)[:-1])), cnt)
    return starts + np.arange(total, dtype=np.int64), cnt


class CudaLoops:
    """One instantiation (float or double), like one of the reference's two compiled modules."""

    def __init__(self, use_float, device=None):
        self.use_float = bool(use_float)
        self.c_real_t = ctypes.c_float if use_float else ctypes.c_double       # cython_float.pxi:9 / cython_double.pxi:7
        self.obj_ind_type = ctypes.c_size_t                                     # cython_*_nonwindows.pyx:10
        self.obj_long_double_type = ctypes.c_longdouble
        self.real_bytes = 4 if use_float else 8
        self.dtype = np.dtype(np.float32 if use_float else np.float64)
        self._device = device
        #: options forwarded to every engine this module creates (see hpf_set_option)
        self.engine_options = {}
        #: telemetry of the last fit_hpf call
        self.last_stats = {}

    # ---- helpers the reference exports (pxi:11-18) -------------------------------------------------
    def cast_real_t(self, n):
        return self.dtype.type(n).item() if n is not None else None

    @staticmethod
    def cast_int(n):
        return int(n)

    @staticmethod
    def cast_ind_type(n):
        return int(n)

    @property
    def device(self):
        if self._device is not None:
            return self._device
        return int(os.environ.get("LOCAL_RANK", "0")) if "HPF_DEVICE" not in os.environ \
            else int(os.environ["HPF_DEVICE"])

    def _engine(self, nU, nI, k):
        eng = Engine(nU, nI, k, self.real_bytes, self.device)
        for name, value in self.engine_options.items():
            eng.set_option(name, value)
        return eng

    # ---- initialize_parameters (pxi:117-143) ---------------------------------------------------------
    def initialize_parameters(self, Theta, Beta, random_seed, a, a_prime, b_prime, c, c_prime, d_prime):
        """Host-side random start.  Must consume numpy's MT19937 stream exactly as the reference does
        (rates first, then shapes; all four centred on a'/c') so that fits are comparable seed for
        seed; drawn on the host because the bit-stream is numpy's."""
        nU, k = Theta.shape
        nI = Beta.shape[0]
        dt = self.dtype
        gen = np.random.Generator(np.random.MT19937(seed=random_seed if random_seed > 0 else None))

        def jitter(center, rows):
            return center + 0.01 * gen.random(size=(rows, k), dtype=dt)

        Gamma_rte, Lambda_rte = jitter(a_prime, nU), jitter(c_prime, nI)
        Gamma_shp, Lambda_shp = jitter(a_prime, nU), jitter(c_prime, nI)
        k_rte = np.full((nU, 1), b_prime, dtype=dt)
        t_rte = np.full((nI, 1), d_prime, dtype=dt)
        Theta[:, :] = Gamma_shp / Gamma_rte
        Beta[:, :] = Lambda_shp / Lambda_rte
        return Gamma_shp, Gamma_rte, Lambda_shp, Lambda_rte, k_rte, t_rte

    # ---- get_csc_data / get_unique_items_batch (pxi:22-42) ---------------------------------------------
    def get_csc_data(self, ix_u, ix_i, Y, nU, nI):
        from scipy.sparse import coo_array
        X = coo_array((Y, (ix_u, ix_i)), shape=(nU, nI)).tocsc()
        return X.indptr.astype(np.int64), X.indices.astype(np.int64), X.data.astype(self.dtype)

    # ---- SVI epochs (pxi:262-377) ---------------------------------------------------------------------------
    def svi_epochs(self, eng, svi, n_epochs, nU, nI, users_per_batch, items_per_batch, step_size=None):
        """`n_epochs` epochs of the reference's minibatch schedule on a loaded engine.  `svi` carries what persists
        between epochs: the PCG64 generator (pxi:207), the in-place shuffled id lists (pxi:277, 329) and the epoch
        counter.  Epoch e is a USER epoch iff both sizes are set and (e+1) % 2 == 0, or only users_per_batch is set
        (pxi:265-273).  Every minibatch is assembled and applied on the device (hpf_step_batch_ids)."""
        dt = self.dtype
        if step_size is None:
            step_size = lambda x: 1 / np.sqrt(x + 2)   # the constructor default (hpfrec/__init__.py:208)
        for _ in range(int(n_epochs)):
            e = svi["epoch"]
            rho = float(dt.type(step_size(e)))                                       # pxi:263
            if users_per_batch > 0 and items_per_batch > 0:
                user_epoch = ((e + 1) % 2) == 0                                      # pxi:265-269
            else:
                user_epoch = users_per_batch > 0
            ids, n, per = (svi["users"], nU, users_per_batch) if user_epoch else (svi["items"], nI, items_per_batch)
            svi["rng"].shuffle(ids)                                                  # pxi:277 / 329
            eng.step_epoch_ids(ids, per, user_epoch, rho)
            svi["epoch"] = e + 1

    # ---- fit_hpf (pxi:147-418) -----------------------------------------------------------------------
    def fit_hpf(self, a, a_prime, b_prime, c, c_prime, d_prime, Y, ix_u, ix_i, Theta, Beta,
                maxiter, stop_crit, check_every, stop_thr, users_per_batch, items_per_batch,
                step_size, sum_exp_trick, st_ix_u, save_folder, random_seed, verbose, nthreads,
                par_sh, has_valset, Yval, ix_u_val, ix_i_val, full_llk, keep_all_objs,
                alloc_full_phi):
        dt = self.dtype
        nU, k = Theta.shape
        nI = Beta.shape[0]
        nY = Y.shape[0]
        Y = np.ascontiguousarray(Y, dtype=dt)
        ix_u = as_index(ix_u)
        ix_i = as_index(ix_i)
        users_per_batch = int(getattr(users_per_batch, "value", users_per_batch))
        items_per_batch = int(getattr(items_per_batch, "value", items_per_batch))
        if has_valset:
            Yval = np.ascontiguousarray(Yval, dtype=dt)
            ix_u_val, ix_i_val = as_index(ix_u_val), as_index(ix_i_val)
            nYv = Yval.shape[0]

        if verbose > 0:
            print("Initializing parameters...")
        state = self.initialize_parameters(Theta, Beta, random_seed, a, a_prime, b_prime, c, c_prime, d_prime)

        eng = self._engine(nU, nI, k)
        try:
            eng.set_hyper(a, a_prime, b_prime, c, c_prime, d_prime)
            eng.load_state(*state)
            del state
            full_updates = (users_per_batch == 0) and (items_per_batch == 0)
            t_ingest = time.time()
            if not full_updates:
                # minibatches are assembled on the device from plain CSR / CSC orderings (one panel)
                eng.set_option("panel_mb", 1e9)
            # the resident triples feed the sweep, the minibatch assembly and the training-llk checks
            eng.load_coo(ix_u, ix_i, Y)
            t_ingest = time.time() - t_ingest

            if items_per_batch > 0 and verbose:
                print("Creating item indices for stochastic optimization...")
            # shuffled id lists persist across epochs; the generator is PCG64 seeded like pxi:207
            svi = dict(rng=np.random.default_rng(seed=random_seed if random_seed > 0 else None),
                       users=np.arange(nU, dtype=np.int64), items=np.arange(nI, dtype=np.int64), epoch=0)
            errs = [0.0, 0.0]
            last_crit = -np.inf
            Theta_prev = None
            if stop_crit == "diff-norm":
                Theta_prev = Theta.copy()

            if verbose > 0:
                print("Initializing optimization procedure...")
            st_time = time.time()

            def metrics():
                """errs[0] = llk criterion, errs[1] = rmse (assess_convergence, pxi:66-79)."""
                if has_valset:
                    o = eng.llk(ix_u_val, ix_i_val, Yval, full_llk)
                    return o[0] - o[2], np.sqrt(o[1] / nYv)
                o = eng.llk_train(full_llk)
                return o[0] - o[3], np.sqrt(o[1] / nY)

            i = -1
            it_done = 0
            while it_done < maxiter:
                if full_updates:
                    # run up to the next convergence check in ONE device call (no host round trips)
                    burst = maxiter - it_done
                    if check_every > 0:
                        burst = min(burst, check_every - (it_done % check_every))
                    eng.step_full(burst)
                    it_done += burst
                else:
                    self.svi_epochs(eng, svi, 1, nU, nI, users_per_batch, items_per_batch, step_size)
                    it_done += 1
                i = it_done - 1

                # ---- assess_convergence (pxi:51-92, 381-394)
                if check_every > 0 and (it_done % check_every) == 0:
                    converged = False
                    if stop_crit == "diff-norm":
                        eng.export_state(Theta=Theta)
                        last_crit = float(np.linalg.norm(Theta - Theta_prev))
                        if verbose:
                            print("Iteration %d | Norm(Theta_{%d} - Theta_{%d}): %.5f"
                                  % (it_done, it_done, it_done - check_every, last_crit))
                        if last_crit < stop_thr:
                            converged = True
                        else:
                            Theta_prev[:, :] = Theta
                    else:
                        errs = list(metrics())
                        if verbose:
                            kind = "val" if has_valset else "train"
                            print(("Iteration %d | " + kind + " llk: %d | " + kind + " rmse: %.4f")
                                  % (it_done, int(errs[0]), errs[1]))
                        if stop_crit != "maxiter":
                            if it_done == check_every:
                                last_crit = errs[0]
                            else:
                                if (1.0 - errs[0] / last_crit) <= stop_thr:
                                    converged = True
                                else:
                                    last_crit = errs[0]
                    if converged:
                        break

            # ---- eval_after_term (pxi:94-113)
            last_llk = None
            if stop_crit in ("diff-norm", "maxiter") and verbose > 0:
                if has_valset:
                    o = eng.llk(ix_u_val, ix_i_val, Yval, full_llk)
                    # the reference subtracts Theta[ix_u_val].sum(0) . Beta[ix_i_val].sum(0) here (pxi:105)
                    eng.export_state(Theta=Theta, Beta=Beta)
                    cross = Theta[ix_u_val.astype(np.int64)].sum(axis=0).dot(Beta[ix_i_val.astype(np.int64)].sum(axis=0))
                    errs = [o[0] - float(cross), np.sqrt(o[1] / nYv)]
                else:
                    errs = list(metrics())
                last_llk = np.longdouble(errs[0])
            end_tm = (time.time() - st_time) / 60
            if verbose:
                print("\n\nOptimization finished")
                print("Final log-likelihood: %d" % int(errs[0]))
                print("Final RMSE: %.4f" % errs[1])
                print("Minutes taken (optimization part): %.1f" % end_tm)
                print("")

            out = eng.export_all()
            Theta[:, :] = out["Theta"]
            Beta[:, :] = out["Beta"]
            self.last_stats = dict(seconds_loop=end_tm * 60, seconds_ingest=t_ingest,
                                   gpu_launches=eng.launch_count, iterations=it_done)
        finally:
            eng.close()

        if save_folder != "":
            if verbose:
                print("Saving final parameters to .csv files...")
            names = ["Theta", "Beta", "Gamma_shp", "Gamma_rte", "Lambda_shp", "Lambda_rte", "kappa_rte", "tau_rte"]
            objs = [Theta, Beta, out["Gamma_shp"], out["Gamma_rte"], out["Lambda_shp"], out["Lambda_rte"],
                    out["k_rte"], out["t_rte"]]
            for nm, ob in zip(names, objs):
                np.savetxt(os.path.join(save_folder, nm), ob, fmt="%.10f", delimiter=",")

        temp = None
        if keep_all_objs:
            temp = (out["Gamma_shp"], out["Gamma_rte"], out["Lambda_shp"], out["Lambda_rte"],
                    out["k_rte"], out["t_rte"])
        return i, temp, last_llk

    # ---- partial_fit (pxi:423-473) -------------------------------------------------------------------
    def partial_fit(self, Y_batch, ix_u_batch, ix_i_batch, Theta, Beta, Gamma_shp, Gamma_rte,
                    Lambda_shp, Lambda_rte, k_rte, t_rte, add_k_rte, add_t_rte, a, c, k_shp, t_shp, k,
                    users_this_batch, items_this_batch, par_sh, step_size_batch, multiplier_batch,
                    nthreads, user_batch):
        """One user-supplied minibatch; all eight arrays are updated IN PLACE like the reference's
        `[:,:] =` assignments."""
        dt = self.dtype
        nU, nI = Gamma_shp.shape[0], Lambda_shp.shape[0]
        eng = self._engine(nU, nI, int(k))
        try:
            eng.set_constants(a, c, k_shp, t_shp, add_k_rte, add_t_rte)
            eng.load_state(np.ascontiguousarray(Gamma_shp, dt), np.ascontiguousarray(Gamma_rte, dt),
                           np.ascontiguousarray(Lambda_shp, dt), np.ascontiguousarray(Lambda_rte, dt),
                           np.ascontiguousarray(k_rte, dt), np.ascontiguousarray(t_rte, dt))
            iu, ii = as_index(ix_u_batch).astype(np.int64), as_index(ix_i_batch).astype(np.int64)
            eng.step_batch(iu, ii, np.ascontiguousarray(Y_batch, dt),
                           as_index(users_this_batch).astype(np.int64), as_index(items_this_batch).astype(np.int64),
                           bool(user_batch), float(dt.type(step_size_batch)), float(dt.type(multiplier_batch)), True)
            out = eng.export_all()
        finally:
            eng.close()
        for dst, key in ((Theta, "Theta"), (Beta, "Beta"), (Gamma_shp, "Gamma_shp"), (Gamma_rte, "Gamma_rte"),
                         (Lambda_shp, "Lambda_shp"), (Lambda_rte, "Lambda_rte"), (k_rte, "k_rte"), (t_rte, "t_rte")):
            dst[...] = out[key].reshape(dst.shape)

    # ---- calc_user_factors (pxi:476-520) ---------------------------------------------------------------
    def calc_user_factors(self, a, a_prime, b_prime, c, c_prime, d_prime, Y, ix_i, Theta, Beta,
                          Lambda_shp, Lambda_rte, nY, k, maxiter, nthreads, random_seed, stop_thr, return_all):
        """Single-user CAVI with the item side frozen: the user side of the full-batch iteration on a
        one-row engine (passes + hpf_update_users; hpf_update_items is never called)."""
        dt = self.dtype
        k = int(k)
        nI = Beta.shape[0]
        k_shp = dt.type(a_prime + k * a)
        rng = np.random.default_rng(seed=random_seed if random_seed > 0 else None)
        Theta[:] = rng.gamma(a, 1 / b_prime, size=k).astype(dt)                       # pxi:491
        k_rte = dt.type(b_prime + Theta.sum())                                         # pxi:492
        Gamma_rte = rng.gamma(a_prime, b_prime / a_prime, size=1).astype(dt) + Beta.sum(axis=0)   # pxi:493
        Gamma_shp = Gamma_rte * Theta * rng.uniform(low=.85, high=1.15, size=k).astype(dt)       # pxi:495
        np.nan_to_num(Gamma_shp, copy=False)
        np.nan_to_num(Gamma_rte, copy=False)
        Gamma_shp = np.ascontiguousarray(Gamma_shp.reshape(1, k), dtype=dt)
        Gamma_rte = np.ascontiguousarray(Gamma_rte.reshape(1, k), dtype=dt)
        Theta_prev = Theta.copy()
        Y = np.ascontiguousarray(Y, dtype=dt)
        ix_i = as_index(ix_i).astype(np.int64)
        eng = self._engine(1, nI, k)
        try:
            eng.set_hyper(a, a_prime, b_prime, c, c_prime, d_prime)
            eng.load_state(Gamma_shp, Gamma_rte, np.ascontiguousarray(Lambda_shp, dt),
                           np.ascontiguousarray(Lambda_rte, dt), np.full((1, 1), k_rte, dtype=dt),
                           np.ones((nI, 1), dtype=dt))
            eng.load_coo(np.zeros(int(nY), dtype=np.int64), ix_i, Y)                  # pxi:501 (all-zero ix_u)
            th = np.empty((1, k), dtype=dt)
            G_read, R_read = Gamma_shp.copy(), Gamma_rte.copy()   # state the last update_phi read (for phi)
            user_pass_only = eng.describe().get("robust") != "1"   # the item side is frozen: its sums are never used
            for _ in range(int(maxiter)):
                if return_all:
                    eng.export_state(Gamma_shp=G_read, Gamma_rte=R_read)
                if user_pass_only:
                    eng.sweep_side(1)
                else:
                    eng.sweep()
                eng.update_users()                                                     # pxi:507-510
                eng.export_state(Theta=th)
                Theta[:] = th[0]
                if np.linalg.norm(Theta - Theta_prev) < stop_thr:                      # pxi:513
                    break
                Theta_prev = Theta.copy()
            if not return_all:
                return None
            eng.export_state(Gamma_shp=Gamma_shp, Gamma_rte=Gamma_rte)
        finally:
            eng.close()
        # phi / Y of the last sweep (pxi:518): the engine never stores phi, so it is re-evaluated from
        # the state that sweep read
        from .engine import update_shapes
        G_tmp, L_tmp = G_read, np.ascontiguousarray(Lambda_shp, dt).copy()
        phi = np.empty((int(nY), k), dtype=dt)
        update_shapes(G_tmp, R_read, L_tmp, np.ascontiguousarray(Lambda_rte, dt), Y,
                      np.zeros(int(nY), dtype=np.int64), ix_i, a, c, phi=phi, device=self.device)
        return Gamma_shp.reshape(-1), Gamma_rte.reshape(-1), phi / Y.reshape((-1, 1))

    def calc_user_factors_batch(self, a, a_prime, b_prime, c, c_prime, d_prime, Y, ix_u, ix_i, Beta, Lambda_shp, Lambda_rte,
                                n_users, k, maxiter, random_seed, stop_thr, return_all=False):
        """calc_user_factors (pxi:476-520) for MANY new users in one engine: `ix_u` numbers the new users 0..n_users-1,
        the item side is frozen.  Every user gets exactly what a separate calc_user_factors call with the same
        `random_seed` gives her (the reference seeds its generator per call, so all users start from the same draw; a
        user's iterations never read another user's rows), including her own early stop: the device keeps iterating
        all rows, the host freezes each user's result at the first iteration where ||Theta_t - Theta_{t-1}|| < stop_thr.
        Per iteration: ONE user-major pass + ONE row update for all users, one (n_users x k) read-back.
        Returns Theta (n_users, k), or (Theta, Gamma_shp, Gamma_rte, n_iter) with return_all."""
        dt = self.dtype
        k, B = int(k), int(n_users)
        nI = Beta.shape[0]
        k_shp = dt.type(a_prime + k * a)
        rng = np.random.default_rng(seed=random_seed if random_seed > 0 else None)
        theta0 = rng.gamma(a, 1 / b_prime, size=k).astype(dt)                              # pxi:491
        k_rte0 = dt.type(b_prime + theta0.sum())                                            # pxi:492
        rte0 = rng.gamma(a_prime, b_prime / a_prime, size=1).astype(dt) + Beta.sum(axis=0)  # pxi:493
        shp0 = rte0 * theta0 * rng.uniform(low=.85, high=1.15, size=k).astype(dt)          # pxi:495
        np.nan_to_num(shp0, copy=False)
        np.nan_to_num(rte0, copy=False)
        Y = np.ascontiguousarray(Y, dtype=dt)
        ix_u, ix_i = as_index(ix_u).astype(np.int64), as_index(ix_i).astype(np.int64)
        if ix_u.shape[0] and (ix_u.min() < 0 or ix_u.max() >= B):
            raise ValueError("user numbers must lie in [0, n_users)")
        Theta = np.tile(theta0.astype(dt), (B, 1))
        out_shp = np.tile(shp0.astype(dt), (B, 1))
        out_rte = np.tile(rte0.astype(dt), (B, 1))
        n_iter = np.zeros(B, dtype=np.int64)
        active = np.ones(B, dtype=bool)
        prev = Theta.copy()
        eng = self._engine(B, nI, k)
        try:
            eng.set_hyper(a, a_prime, b_prime, c, c_prime, d_prime)
            eng.load_state(np.ascontiguousarray(out_shp), np.ascontiguousarray(out_rte), np.ascontiguousarray(Lambda_shp, dt),
                           np.ascontiguousarray(Lambda_rte, dt), np.full((B, 1), k_rte0, dtype=dt), np.ones((nI, 1), dtype=dt))
            eng.load_coo(ix_u, ix_i, Y)
            user_pass_only = eng.describe().get("robust") != "1"
            th = np.empty((B, k), dtype=dt)
            gs = np.empty((B, k), dtype=dt) if return_all else None
            gr = np.empty((B, k), dtype=dt) if return_all else None
            for _ in range(int(maxiter)):
                if user_pass_only:
                    eng.sweep_side(1)
                else:
                    eng.sweep()
                eng.update_users()                                                          # pxi:507-510, all users at once
                if return_all:
                    eng.export_state(Gamma_shp=gs, Gamma_rte=gr, Theta=th)
                else:
                    eng.export_state(Theta=th)
                Theta[active] = th[active]
                n_iter[active] += 1
                if return_all:
                    out_shp[active], out_rte[active] = gs[active], gr[active]
                done = active & (np.linalg.norm(th - prev, axis=1) < stop_thr)              # pxi:513, per user
                active &= ~done
                if not active.any():
                    break
                prev[active] = th[active]
        finally:
            eng.close()
        if return_all:
            return Theta, out_shp, out_rte, n_iter
        return Theta

    # ---- calc_llk (pxi:525-534) / predict_arr (pxi:538-543) --------------------------------------------
    def _factors_engine(self, Theta, Beta):
        dt = self.dtype
        nU, k = Theta.shape
        nI = Beta.shape[0]
        eng = self._engine(nU, nI, k)
        eng.load_state(np.ascontiguousarray(Theta, dt), np.ones((nU, k), dt), np.ascontiguousarray(Beta, dt),
                       np.ones((nI, k), dt), np.ones((nU, 1), dt), np.ones((nI, 1), dt))
        return eng

    def calc_llk(self, Y, ix_u, ix_i, Theta, Beta, k, nthreads, full_llk):
        eng = self._factors_engine(Theta, Beta)
        try:
            o = eng.llk(as_index(ix_u), as_index(ix_i), np.ascontiguousarray(Y, self.dtype), full_llk)
        finally:
            eng.close()
        return np.longdouble(o[0]) - np.longdouble(o[2])

    def predict_arr(self, M1, M2, ix_u, ix_i, nthreads):
        eng = self._factors_engine(M1, M2)
        try:
            return eng.predict(as_index(ix_u), as_index(ix_i))
        finally:
            eng.close()


cuda_loops_float = CudaLoops(True)
cuda_loops_double = CudaLoops(False)
