"""hpfrec_b200 -- Hierarchical Poisson Factorization with the variational sweep on NVIDIA B200.

`HPF` keeps the constructor arguments, methods, attributes and (UserId, ItemId, Count) input of
david-cortes/hpfrec's `hpfrec.HPF` (reference hpfrec/__init__.py:11-1458) so existing scripts run
unchanged, but every numeric loop of fitting / scoring runs as sm_100a CUDA kernels through
libhpf_b200.so (see hpfrec_b200/loops.py for the module-level mirror of the reference's compiled
extension and include/hpf_b200.h for the C ABI).  No CPU fallback exists: without the CUDA library or
without a GPU, fitting raises.
"""
import inspect
import multiprocessing
import os
import types
import warnings

import numpy as np
import pandas as pd
from scipy.sparse import coo_array, issparse

from .loops import CudaLoops, cuda_loops_double, cuda_loops_float
from .engine import Engine

__all__ = ["HPF", "Engine", "cuda_loops_float", "cuda_loops_double"]
__version__ = "0.1.0"

_STOP_CRITERIA = ("maxiter", "train-llk", "val-llk", "diff-norm")
_COLS = ["UserId", "ItemId", "Count"]


def _as_positive_float(name, value):
    if isinstance(value, int):
        value = float(value)
    assert isinstance(value, float), "'%s' must be a float" % name
    assert value > 0, "'%s' must be positive" % name
    return value


def _batch_size(value):
    if value is None:
        return 0
    if isinstance(value, float):
        value = int(value)
    assert isinstance(value, int)
    assert value > 0
    return value


def _triplets_frame(data, what, copy=True):
    """ndarray (first three columns) / DataFrame -> DataFrame with exactly UserId, ItemId, Count."""
    if isinstance(data, np.ndarray):
        assert len(data.shape) > 1
        assert data.shape[1] >= 3
        return pd.DataFrame(data[:, :3], copy=copy, columns=_COLS)
    if isinstance(data, pd.DataFrame):
        assert data.shape[0] > 0
        for col in _COLS:
            assert col in data.columns
        return data[_COLS].copy()
    return None


def _codes(values, mapping):
    """Row numbers of `values` in `mapping` (-1 where absent)."""
    return np.require(pd.Categorical(values, mapping).codes, requirements=["ENSUREARRAY"])


class HPF:
    """Hierarchical Poisson Factorization (Gopalan, Hofman & Blei 2015) fitted by mean-field
    coordinate ascent (full batch) or stochastic variational inference (minibatches of users and/or
    items) on the GPU.

    Parameters are those of `hpfrec.HPF` (reference hpfrec/__init__.py:87-177), with identical names,
    defaults and validation: k, a, a_prime, b_prime, c, c_prime, d_prime, ncores, stop_crit
    ('maxiter' | 'train-llk' | 'val-llk' | 'diff-norm'), check_every, stop_thr, users_per_batch,
    items_per_batch, step_size, maxiter, use_float, reindex, verbose, random_seed,
    allow_inconsistent_math, full_llk, alloc_full_phi, keep_data, save_folder, produce_dicts,
    keep_all_objs, sum_exp_trick.

    `ncores`, `allow_inconsistent_math`, `alloc_full_phi` and `sum_exp_trick` steer the reference's CPU
    loops only; they are validated and stored but do not change the GPU computation (which is always
    race-free, never materialises phi and always normalises before exponentiating).

    Attributes after fitting: Theta (nusers, k), Beta (nitems, k), user_mapping_, item_mapping_,
    user_dict_, item_dict_, is_fitted, niter, train_llk and -- with keep_all_objs -- Gamma_shp,
    Gamma_rte, Lambda_shp, Lambda_rte, k_rte, t_rte.
    """

    def __init__(self, k=30, a=0.3, a_prime=0.3, b_prime=1.0,
                 c=0.3, c_prime=0.3, d_prime=1.0, ncores=-1,
                 stop_crit='maxiter', check_every=10, stop_thr=1e-3,
                 users_per_batch=None, items_per_batch=None, step_size=lambda x: 1 / np.sqrt(x + 2),
                 maxiter=100, use_float=True, reindex=True, verbose=True,
                 random_seed=None, allow_inconsistent_math=False, full_llk=False,
                 alloc_full_phi=False, keep_data=True, save_folder=None,
                 produce_dicts=True, keep_all_objs=True, sum_exp_trick=False):
        assert isinstance(k, int)
        assert k > 0
        self.k = k
        self.a = _as_positive_float("a", a)
        self.a_prime = _as_positive_float("a_prime", a_prime)
        self.b_prime = _as_positive_float("b_prime", b_prime)
        self.c = _as_positive_float("c", c)
        self.c_prime = _as_positive_float("c_prime", c_prime)
        self.d_prime = _as_positive_float("d_prime", d_prime)

        if ncores is None:
            ncores = 1
        elif ncores < 1:
            ncores = multiprocessing.cpu_count()
        assert isinstance(ncores, int) and ncores > 0
        self.ncores = ncores

        if random_seed is not None:
            assert isinstance(random_seed, int)
        assert stop_crit in _STOP_CRITERIA

        if maxiter is None:
            if stop_crit == 'maxiter':
                raise ValueError("If 'stop_crit' is set to 'maxiter', must provide a maximum number of iterations.")
            maxiter = 10 ** 10
        else:
            assert isinstance(maxiter, int)
            assert maxiter > 0

        if check_every is None:
            if stop_crit != 'maxiter':
                raise ValueError("If 'stop_crit' is not 'maxiter', must input after how many iterations to calculate it.")
            check_every = 0
        else:
            assert isinstance(check_every, int)
            assert check_every > 0
            assert check_every <= maxiter

        if isinstance(stop_thr, int):
            stop_thr = float(stop_thr)
        if stop_thr is not None:
            assert isinstance(stop_thr, float)
            assert stop_thr > 0

        if save_folder is not None:
            save_folder = os.path.expanduser(save_folder)
            assert os.path.exists(save_folder)

        verbose = bool(verbose)
        if stop_crit == 'maxiter' and not verbose:
            check_every = 0          # nothing would consume the metric

        if not isinstance(step_size, types.FunctionType):
            raise ValueError("'step_size' must be a function.")
        if len(inspect.getfullargspec(step_size).args) < 1:
            raise ValueError("'step_size' must be able to take the iteration number as input.")
        for probe in (0, 1):
            assert 0 <= step_size(probe) <= 1

        self.users_per_batch = _batch_size(users_per_batch)
        self.items_per_batch = _batch_size(items_per_batch)
        self.step_size = step_size
        self.allow_inconsistent_math = bool(allow_inconsistent_math)
        self.use_float = bool(use_float)
        # engine limit (the reference has none): a factor row must fit the widest lane-group shape, 128 packs of 16 bytes
        kmax = 512 if self.use_float else 256
        if self.k > kmax:
            raise ValueError("k=%d exceeds what the CUDA engine supports for %s (k <= %d)"
                             % (self.k, "float32" if self.use_float else "float64", kmax))
        self.random_seed = random_seed
        self.stop_crit = stop_crit
        self.reindex = bool(reindex)
        self.keep_data = bool(keep_data)
        self.maxiter = maxiter
        self.check_every = check_every
        self.stop_thr = stop_thr
        self.save_folder = save_folder
        self.verbose = verbose
        self.produce_dicts = bool(produce_dicts) and self.reindex
        self.full_llk = bool(full_llk)
        self.alloc_full_phi = bool(alloc_full_phi)
        self.keep_all_objs = bool(keep_all_objs)
        self.sum_exp_trick = bool(sum_exp_trick)

        self.Theta = None
        self.Beta = None
        self._scorer_obj, self._scorer_key = None, None
        self.user_mapping_ = None
        self.item_mapping_ = None
        self.user_dict_ = None
        self.item_dict_ = None
        self.is_fitted = False
        self.niter = None
        self.train_llk = None

    # ------------------------------------------------------------------------------------------------
    @property
    def _loops(self) -> CudaLoops:
        """The float or double instantiation of the GPU hot path (the reference picks between its two
        compiled modules the same way at every call site, e.g. hpfrec/__init__.py:508)."""
        return cuda_loops_float if self.use_float else cuda_loops_double

    # ---- device-resident factors for predict / topN / eval_llk ---------------------------------------
    def _scorer(self):
        """Theta and Beta on the device, uploaded once per fitted state (they are re-uploaded only after a call
        that changes them: fit, partial_fit, add_user).  The reference reads its host arrays directly
        (hpfrec/__init__.py:1277-1291, 1338-1396, 1433); here the same reads are device gathers."""
        from .engine import Scorer
        key = (id(self.Theta), id(self.Beta), self.Theta.shape, self.Beta.shape)
        if getattr(self, "_scorer_obj", None) is None or getattr(self, "_scorer_key", None) != key:
            self._drop_scorer()
            self._scorer_obj = Scorer(self.Theta, self.Beta, device=self._loops.device)
            self._scorer_key = key
        return self._scorer_obj

    def _drop_scorer(self):
        obj = getattr(self, "_scorer_obj", None)
        if obj is not None:
            obj.close()
        self._scorer_obj, self._scorer_key = None, None

    def __getstate__(self):   # the device handle does not pickle (the reference pickles with dill, README.md:162-173)
        state = dict(self.__dict__)
        state["_scorer_obj"], state["_scorer_key"] = None, None
        return state

    def _typed(self, frame):
        """Count -> real_t, ids -> the reference's index type (reference __init__.py:508-514)."""
        lp = self._loops
        if frame["Count"].dtype != lp.c_real_t:
            frame["Count"] = frame["Count"].astype(lp.c_real_t)
        for col in ("UserId", "ItemId"):
            if frame[col].dtype != lp.obj_ind_type:
                frame[col] = frame[col].astype(lp.obj_ind_type)
        return frame

    # ------------------------------------------------------------------------------------------------
    def fit(self, counts_df, val_set=None):
        """Fits the model to (UserId, ItemId, Count) triplets given as a DataFrame, an array with
        those three columns, or a scipy COO array (which forces reindex=False).  `val_set` (same
        format) is only used with stop_crit='val-llk' / 'maxiter'.  Returns self.
        Inputs may be modified in place, as in the reference."""
        if self.stop_crit == 'val-llk' and val_set is None:
            raise ValueError("If 'stop_crit' is set to 'val-llk', must provide a validation set.")
        if self.verbose:
            self._print_st_msg()
        self._process_data(counts_df)
        if self.verbose:
            self._print_data_info()
        if (val_set is not None) and (self.stop_crit not in ("diff-norm", "train-llk")):
            self._process_valset(val_set)
        else:
            self.val_set = None

        self._cast_before_fit()
        self._fit()

        if self.keep_data:
            if self.users_per_batch == 0:
                self._store_metadata()
            else:
                self._st_ix_user = self._st_ix_user[:-1]
        if self.produce_dicts and self.reindex:
            self.user_dict_ = {self.user_mapping_[i]: i for i in range(self.user_mapping_.shape[0])}
            self.item_dict_ = {self.item_mapping_[i]: i for i in range(self.item_mapping_.shape[0])}
        self.is_fitted = True
        self._drop_scorer()
        del self.input_df
        del self.val_set
        return self

    def _process_data(self, input_df):
        sizes_known = False
        frame = _triplets_frame(input_df, "counts_df")
        if frame is None:
            if issparse(input_df) and input_df.format == "coo":
                self.nusers, self.nitems = input_df.shape
                frame = pd.DataFrame({"UserId": input_df.row, "ItemId": input_df.col,
                                      "Count": input_df.data}, copy=False)
                self.reindex = False
                sizes_known = True
            else:
                raise ValueError("'input_df' must be a pandas data frame, numpy array, or scipy sparse coo_array.")

        cutoff = 0 if self.stop_crit in ('maxiter', 'diff-norm') else 0.9
        too_small = frame["Count"] <= cutoff
        if too_small.sum() > 0:
            warnings.warn(
                "'counts_df' contains observations with a count value less than 1, these will be ignored."
                " Any user or item associated exclusively with zero-value observations will be excluded."
                " If using 'reindex=False', make sure that your data still meets the necessary criteria."
                " If you still want to use these observations, set 'stop_crit' to 'diff-norm' or 'maxiter'.")
            frame = frame.loc[~too_small]
        self.input_df = frame

        if self.reindex:
            frame["UserId"], umap = self._factorize(frame["UserId"])
            frame["ItemId"], imap = self._factorize(frame["ItemId"])
            self.user_mapping_ = np.require(umap, requirements=["ENSUREARRAY"]).reshape(-1)
            self.item_mapping_ = np.require(imap, requirements=["ENSUREARRAY"]).reshape(-1)
            self.nusers = self.user_mapping_.shape[0]
            self.nitems = self.item_mapping_.shape[0]
            if self.save_folder is not None:
                if self.verbose:
                    print("\nSaving user and item mappings...\n")
                pd.Series(self.user_mapping_).to_csv(os.path.join(self.save_folder, 'users.csv'), index=False)
                pd.Series(self.item_mapping_).to_csv(os.path.join(self.save_folder, 'items.csv'), index=False)
        elif not sizes_known:
            self.nusers = int(frame["UserId"].max()) + 1
            self.nitems = int(frame["ItemId"].max()) + 1

        if self.save_folder is not None:
            lines = ["%s: %.3f\n" % (nm, getattr(self, nm))
                     for nm in ("a", "a_prime", "b_prime", "c", "c_prime", "d_prime")]
            lines.append("k: %d\n" % self.k)
            lines.append("random seed: %d\n" % self.random_seed if self.random_seed is not None
                         else "random seed: None\n")
            with open(os.path.join(self.save_folder, "hyperparameters.txt"), "w") as pf:
                pf.writelines(lines)

        self._typed(frame)

        if self.users_per_batch != 0:
            if self.nusers < self.users_per_batch:
                warnings.warn("Batch size passed is larger than number of users. Will set it to nusers/10.")
                self.users_per_batch = int(np.ceil(self.nusers / 10))
            # the reference sorts the frame by user here (hpfrec/__init__.py:520) because its loops index the arrays
            # through st_ix_u; the engine builds its own orderings on the device from unsorted triples, so the
            # 48M-row host sort is skipped
            self._store_metadata(for_partial_fit=True)
        return None

    def _factorize(self, column):
        """pd.factorize (hpfrec/__init__.py:478-479).  Integer id columns are factorized on the device (two radix
        sorts, hpf_factorize); anything else (strings, objects, floats) has no device representation and stays with
        pandas.  Same result either way: codes in order of first appearance."""
        values = column.to_numpy(copy=False) if hasattr(column, "to_numpy") else np.asarray(column)
        if values.dtype.kind in "iu" and values.dtype.itemsize in (4, 8) and values.shape[0] > 0:
            from .engine import factorize
            return factorize(values, device=self._loops.device)
        return pd.factorize(column)

    def _process_valset(self, val_set, valset=True):
        frame = _triplets_frame(val_set, "val_set")
        if frame is None:
            if issparse(val_set) and val_set.format == "coo":
                assert val_set.shape[0] <= self.nusers
                assert val_set.shape[1] <= self.nitems
                frame = pd.DataFrame({"UserId": val_set.row, "ItemId": val_set.col, "Count": val_set.data},
                                     copy=False)
            else:
                raise ValueError("'val_set' must be a pandas data frame, numpy array, or sparse coo_array.")
        self.val_set = frame

        cutoff = 0 if self.stop_crit == 'val-llk' else 0.9
        too_small = frame["Count"] <= cutoff
        if too_small.sum() > 0:
            warnings.warn("'val_set' contains observations with a count value less than 1, these will be ignored.")
            self.val_set = frame = frame.loc[~too_small]

        if self.reindex:
            frame['UserId'] = _codes(frame["UserId"], self.user_mapping_)
            frame['ItemId'] = _codes(frame["ItemId"], self.item_mapping_)
            self.val_set = frame = frame.loc[(frame["UserId"] != -1) & (frame["ItemId"] != -1)]
            if frame.shape[0] == 0:
                if not valset:
                    raise ValueError("'input_df' has no combinations of users and items"
                                     "in common with the training set.")
                warnings.warn("Validation set has no combinations of users and items"
                              " in common with training set. If 'stop_crit' was set"
                              " to 'val-llk', will now be switched to 'train-llk'.")
                if self.stop_crit == 'val-llk':
                    self.stop_crit = 'train-llk'
                self.val_set = None
                return None
            frame.reset_index(drop=True, inplace=True)
        self._typed(self.val_set)
        return None

    def _store_metadata(self, for_partial_fit=False):
        if self.verbose and for_partial_fit:
            print("Creating user indices for stochastic optimization...")
        # the reference builds scipy's coo -> csr here (hpfrec/__init__.py:591-598) only for its indptr / indices; the
        # same two arrays come from one device sort of the (user, item) pairs (hpf_csr_metadata)
        from .engine import csr_metadata
        df = self.input_df
        indptr, indices = csr_metadata(df["UserId"].to_numpy(copy=False), df["ItemId"].to_numpy(copy=False),
                                       self.nusers, self.nitems, device=self._loops.device)
        self._n_seen_by_user = indptr[1:] - indptr[:-1]
        if for_partial_fit:
            self._st_ix_user = np.require(indptr, dtype=self._loops.obj_ind_type,
                                          requirements=["ENSUREARRAY", "C_CONTIGUOUS"])
        else:
            self._st_ix_user = indptr[:-1]
        self.seen = indices
        return None

    def _cast_before_fit(self):
        lp = self._loops
        self.nusers = int(self.nusers)
        self.nitems = int(self.nitems)
        self.Theta = np.empty((self.nusers, self.k), dtype=lp.c_real_t)
        self.Beta = np.empty((self.nitems, self.k), dtype=lp.c_real_t)
        self.verbose = int(self.verbose)
        if self.random_seed is None:
            self.random_seed = 0
        self.stop_thr = lp.cast_real_t(self.stop_thr)
        for nm in ("a", "a_prime", "b_prime", "c", "c_prime", "d_prime"):
            setattr(self, nm, lp.cast_real_t(getattr(self, nm)))
        if self.save_folder is None:
            self.save_folder = ""

    def _fit(self):
        lp = self._loops
        real, ind = lp.c_real_t, lp.obj_ind_type

        def col(frame, name, dtype):
            return np.require(frame[name].to_numpy(copy=False), dtype=dtype,
                              requirements=["ENSUREARRAY", "C_CONTIGUOUS"])

        has_val = self.val_set is not None
        if has_val:
            yv, uv, iv = col(self.val_set, "Count", real), col(self.val_set, "UserId", ind), col(self.val_set, "ItemId", ind)
        else:
            yv, uv, iv = np.empty(0, dtype=real), np.empty(0, dtype=ind), np.empty(0, dtype=ind)
        if self.users_per_batch == 0:
            self._st_ix_user = np.arange(1).astype(ind)

        self.niter, temp, self.train_llk = lp.fit_hpf(
            self.a, self.a_prime, self.b_prime, self.c, self.c_prime, self.d_prime,
            col(self.input_df, "Count", real), col(self.input_df, "UserId", ind), col(self.input_df, "ItemId", ind),
            self.Theta, self.Beta,
            int(self.maxiter), self.stop_crit, int(self.check_every), self.stop_thr,
            self.users_per_batch, self.items_per_batch,
            self.step_size, int(self.sum_exp_trick),
            self._st_ix_user.astype(ind),
            self.save_folder, int(self.random_seed), self.verbose,
            self.ncores, int(self.allow_inconsistent_math),
            int(has_val), yv, uv, iv,
            int(self.full_llk), int(self.keep_all_objs), int(self.alloc_full_phi))

        if self.users_per_batch == 0:
            del self._st_ix_user
        if self.keep_all_objs:
            (self.Gamma_shp, self.Gamma_rte, self.Lambda_shp, self.Lambda_rte, self.k_rte, self.t_rte) = temp

    # ------------------------------------------------------------------------------------------------
    def partial_fit(self, counts_df, batch_type='users', step_size=None,
                    nusers=None, nitems=None, users_in_batch=None, items_in_batch=None,
                    new_users=False, new_items=False, random_seed=None):
        """Updates the model with ALL the non-zero entries of a subset of users (batch_type='users')
        or of items ('items').  Requires reindex=False and keep_all_objs=True; ids must be
        0..n-1.  On a never-fitted object pass the totals `nusers` and `nitems`.  Semantics follow
        reference hpfrec/__init__.py:714-931 and cython_loops.pxi:423-473."""
        if self.reindex:
            raise ValueError("'partial_fit' can only be called when using reindex=False.")
        if not self.keep_all_objs:
            raise ValueError("'partial_fit' can only be called when using keep_all_objs=True.")
        if self.keep_data:
            if hasattr(self, "seen"):
                warnings.warn("When using 'partial_fit', the list of items seen by each user is not updated "
                              "with the data passed here.")
            else:
                warnings.warn("When fitting the model through 'partial_fit' without calling 'fit' beforehand, "
                              "'keep_data' will be forced to False.")
                self.keep_data = False

        assert batch_type in ('users', 'items')
        user_batch = batch_type == 'users'

        if nusers is None:
            nusers = getattr(self, "nusers", None)
            if nusers is None:
                raise ValueError("Must specify total number of users when calling 'partial_fit' for the first time.")
        if nitems is None:
            nitems = getattr(self, "nitems", None)
            if nitems is None:
                raise ValueError("Must specify total number of items when calling 'partial_fit' for the first time.")
        if getattr(self, "nusers", None) is None:
            self.nusers = nusers
        if getattr(self, "nitems", None) is None:
            self.nitems = nitems

        if step_size is None:
            if self.niter is None:
                self.niter = 0
                step_size = 1.0
            else:
                try:
                    step_size = self.step_size(self.niter)
                except Exception:
                    step_size = 1 / np.sqrt(self.niter + 2)
        assert 0 <= step_size <= 1

        if random_seed is not None:
            if isinstance(random_seed, float):
                random_seed = int(random_seed)
            assert isinstance(random_seed, int)

        if isinstance(counts_df, np.ndarray):
            counts_df = pd.DataFrame(counts_df[:, :3], copy=False, columns=_COLS)
        assert isinstance(counts_df, pd.DataFrame)
        for name in _COLS:
            assert name in counts_df.columns
        assert counts_df.shape[0] > 0

        lp = self._loops
        req = ["ENSUREARRAY", "C_CONTIGUOUS"]
        Y_batch = np.require(counts_df["Count"], dtype=lp.c_real_t, requirements=req)
        ix_u_batch = np.require(counts_df["UserId"], dtype=lp.obj_ind_type, requirements=req)
        ix_i_batch = np.require(counts_df["ItemId"], dtype=lp.obj_ind_type, requirements=req)
        users_in_batch = np.unique(ix_u_batch) if users_in_batch is None else \
            np.require(users_in_batch, dtype=lp.obj_ind_type, requirements=req)
        items_in_batch = np.unique(ix_i_batch) if items_in_batch is None else \
            np.require(items_in_batch, dtype=lp.obj_ind_type, requirements=req)

        if (self.Theta is None) or (self.Beta is None):
            self._cast_before_fit()
            (self.Gamma_shp, self.Gamma_rte, self.Lambda_shp, self.Lambda_rte, self.k_rte, self.t_rte) = \
                lp.initialize_parameters(self.Theta, self.Beta, self.random_seed, self.a, self.a_prime,
                                         self.b_prime, self.c, self.c_prime, self.d_prime)

        if new_users:
            n_add = self.nusers - (ix_u_batch.max() + 1)
            if n_add < 1:
                raise ValueError("There are no new users in the data passed to 'partial_fit'.")
            self._initialize_extra_users(int(n_add), random_seed)
            self.nusers += n_add
        if new_items:
            n_add = self.nitems - (ix_i_batch.max() + 1)
            if n_add < 1:
                raise ValueError("There are no new items in the data passed to 'partial_fit'.")
            self._initialize_extra_items(int(n_add), random_seed)
            self.nitems += n_add

        lp.partial_fit(
            Y_batch, ix_u_batch, ix_i_batch,
            self.Theta, self.Beta, self.Gamma_shp, self.Gamma_rte, self.Lambda_shp, self.Lambda_rte,
            self.k_rte, self.t_rte,
            lp.cast_real_t(self.a_prime / self.b_prime), lp.cast_real_t(self.c_prime / self.d_prime),
            self.a, self.c,
            lp.cast_real_t(self.a_prime + self.k * self.a), lp.cast_real_t(self.c_prime + self.k * self.c),
            int(self.k), users_in_batch, items_in_batch, int(self.allow_inconsistent_math),
            lp.cast_real_t(step_size), lp.cast_real_t(float(nusers) / users_in_batch.shape[0]),
            self.ncores, user_batch)

        self.niter += 1
        self.is_fitted = True
        self._drop_scorer()
        return self

    def _extra_rows(self, n, seed, center, rate_value):
        """Fresh rows for ids beyond the fitted range (reference __init__.py:933-963): shape, rate,
        expectation, and the hierarchical rate filled with its prior value."""
        dt = self._loops.c_real_t
        rng = np.random.default_rng(seed=seed if (seed is not None and seed > 0) else None)
        shp = center + 0.01 * rng.random(size=(n, self.k), dtype=dt)
        rte = center + 0.01 * rng.random(size=(n, self.k), dtype=dt)
        return shp, rte, shp / rte, np.full((n, 1), rate_value, dtype=dt)

    def _initialize_extra_users(self, n, seed):
        shp, rte, mean, rate = self._extra_rows(n, seed, self.a_prime, self.b_prime)
        self.k_rte = np.r_[self.k_rte, rate]
        self.Theta = np.r_[self.Theta, mean]
        self.Gamma_rte = np.r_[self.Gamma_rte, rte]
        self.Gamma_shp = np.r_[self.Gamma_shp, shp]

    def _initialize_extra_items(self, n, seed):
        shp, rte, mean, rate = self._extra_rows(n, seed, self.c_prime, self.d_prime)
        self.t_rte = np.r_[self.t_rte, rate]
        self.Beta = np.r_[self.Beta, mean]
        self.Lambda_rte = np.r_[self.Lambda_rte, rte]
        self.Lambda_shp = np.r_[self.Lambda_shp, shp]

    # ------------------------------------------------------------------------------------------------
    def _process_data_single(self, counts_df):
        assert self.is_fitted
        assert self.keep_all_objs
        if isinstance(counts_df, np.ndarray):
            assert len(counts_df.shape) > 1
            assert counts_df.shape[1] >= 2
            counts_df = pd.DataFrame(counts_df[:, :2], columns=["ItemId", "Count"], copy=True)
        elif isinstance(counts_df, pd.DataFrame):
            assert counts_df.shape[0] > 0
            assert "ItemId" in counts_df.columns
            assert "Count" in counts_df.columns
            counts_df = counts_df[["ItemId", "Count"]].copy()
        else:
            raise ValueError("'counts_df' must be a pandas data frame or a numpy array")

        if self.reindex:
            if self.produce_dicts:
                try:
                    counts_df["ItemId"] = counts_df["ItemId"].map(lambda x: self.item_dict_[x])
                except Exception:
                    raise ValueError("Can only make calculations for items that were in the training set.")
            else:
                counts_df["ItemId"] = _codes(counts_df["ItemId"].to_numpy(copy=False), self.item_mapping_)
                if (counts_df["ItemId"] == -1).sum() > 0:
                    raise ValueError("Can only make calculations for items that were in the training set.")
        lp = self._loops
        counts_df["ItemId"] = np.require(counts_df["ItemId"], dtype=lp.obj_ind_type)
        counts_df["Count"] = np.require(counts_df["Count"], dtype=lp.c_real_t)
        return counts_df

    @staticmethod
    def _check_input_predict_factors(ncores, random_seed, stop_thr, maxiter):
        if ncores is None:
            ncores = 1
        elif ncores < 1:
            ncores = multiprocessing.cpu_count()
        assert isinstance(ncores, int) and ncores > 0
        assert isinstance(random_seed, int)
        assert random_seed > 0
        if isinstance(stop_thr, int):
            stop_thr = float(stop_thr)
        assert isinstance(stop_thr, float) and stop_thr > 0
        if isinstance(maxiter, float):
            maxiter = int(maxiter)
        assert isinstance(maxiter, int) and maxiter > 0
        return ncores, random_seed, stop_thr, maxiter

    def _user_factors(self, counts_df, maxiter, ncores, random_seed, stop_thr, return_all):
        lp = self._loops
        req = ["ENSUREARRAY", "C_CONTIGUOUS"]
        Theta = np.empty(self.k, dtype=lp.c_real_t)
        temp = lp.calc_user_factors(
            self.a, self.a_prime, self.b_prime, self.c, self.c_prime, self.d_prime,
            np.require(counts_df["Count"].to_numpy(copy=False), dtype=lp.c_real_t, requirements=req),
            np.require(counts_df["ItemId"].to_numpy(copy=False), dtype=lp.obj_ind_type, requirements=req),
            Theta, self.Beta, self.Lambda_shp, self.Lambda_rte,
            int(counts_df.shape[0]), int(self.k), int(maxiter), int(ncores), int(random_seed),
            lp.cast_real_t(stop_thr), int(bool(return_all)))
        if np.isnan(Theta).sum() > 0:
            raise ValueError("NaNs encountered in the result. Failed to produce latent factors.")
        return Theta, temp

    def predict_factors(self, counts_df, maxiter=10, ncores=1, random_seed=1, stop_thr=1e-3, return_all=False):
        """Latent factors of ONE user from her (ItemId, Count) data with the item side frozen
        (reference hpfrec/__init__.py:989-1058).  Returns Theta (k,), or with return_all
        (Theta, Gamma_shp, Gamma_rte, Phi)."""
        ncores, random_seed, stop_thr, maxiter = self._check_input_predict_factors(ncores, random_seed, stop_thr, maxiter)
        counts_df = self._process_data_single(counts_df)
        Theta, temp = self._user_factors(counts_df, maxiter, ncores, random_seed, stop_thr, return_all)
        if return_all:
            return (Theta, temp[0], temp[1], temp[2])
        return Theta

    def predict_factors_batch(self, counts_df, maxiter=10, random_seed=1, stop_thr=1e-3, return_all=False):
        """predict_factors for MANY new users at once (an extension; the reference folds in one user per call,
        hpfrec/__init__.py:989-1058).  `counts_df`: (UserId, ItemId, Count) rows of users that need not be in the model;
        items must be.  Every user receives exactly the factors a separate predict_factors call with the same seed
        returns, but all users share one device engine: one user-major pass and one row update per iteration for the
        whole batch.  Returns (user_ids, Theta) -- user_ids in order of first appearance, Theta (n_users, k) -- or with
        return_all (user_ids, Theta, Gamma_shp, Gamma_rte)."""
        _, random_seed, stop_thr, maxiter = self._check_input_predict_factors(1, random_seed, stop_thr, maxiter)
        assert self.is_fitted and self.keep_all_objs
        if isinstance(counts_df, np.ndarray):
            assert counts_df.ndim == 2 and counts_df.shape[1] >= 3
            counts_df = pd.DataFrame(counts_df[:, :3], columns=["UserId", "ItemId", "Count"])
        assert isinstance(counts_df, pd.DataFrame) and counts_df.shape[0] > 0
        for col in ("UserId", "ItemId", "Count"):
            assert col in counts_df.columns
        codes, user_ids = pd.factorize(counts_df["UserId"].to_numpy(copy=False))
        items = self._process_data_single(counts_df[["ItemId", "Count"]])
        lp = self._loops
        out = lp.calc_user_factors_batch(
            self.a, self.a_prime, self.b_prime, self.c, self.c_prime, self.d_prime,
            items["Count"].to_numpy(copy=False), codes.astype(np.int64), items["ItemId"].to_numpy(copy=False),
            self.Beta, self.Lambda_shp, self.Lambda_rte, int(user_ids.shape[0]), int(self.k), int(maxiter),
            int(random_seed), lp.cast_real_t(stop_thr), return_all=bool(return_all))
        Theta = out[0] if return_all else out
        if np.isnan(Theta).sum() > 0:
            raise ValueError("NaNs encountered in the result. Failed to produce latent factors.")
        if return_all:
            return user_ids, Theta, out[1], out[2]
        return user_ids, Theta

    def add_user(self, user_id, counts_df, update_existing=False, maxiter=10, ncores=1,
                 random_seed=1, stop_thr=1e-3, update_all_params=None):
        """Adds a new user (or, with update_existing, refreshes an existing one) from ALL her
        (ItemId, Count) data without touching item parameters, unless update_all_params is set, in
        which case repeated partial_fit calls are used (reference hpfrec/__init__.py:1060-1196)."""
        ncores, random_seed, stop_thr, maxiter = self._check_input_predict_factors(ncores, random_seed, stop_thr, maxiter)
        if update_existing and self.reindex:
            if self.produce_dicts:
                user_id = self.user_dict_[user_id]
            else:
                user_id = _codes(np.array([user_id]), self.user_mapping_)[0]
                if user_id == -1:
                    raise ValueError("User was not present in the training data.")

        counts_df = self._process_data_single(counts_df)
        lp = self._loops
        if update_all_params:
            counts_df['UserId'] = user_id
            counts_df['UserId'] = np.require(counts_df["UserId"], dtype=lp.obj_ind_type)
            self.partial_fit(counts_df, new_users=(not update_existing))
            Theta_prev = self.Theta[-1].copy()
            for _ in range(maxiter - 1):
                self.partial_fit(counts_df)
                if np.linalg.norm(self.Theta[-1] - Theta_prev) <= stop_thr:
                    break
                Theta_prev = self.Theta[-1].copy()
        else:
            # drop-in parity: the reference passes cast_int(stop_thr) here (hpfrec/__init__.py:1153), i.e. 0, so this
            # path never stops early and always runs `maxiter` iterations (unlike predict_factors, init:1047)
            Theta, temp = self._user_factors(counts_df, maxiter, ncores, random_seed, 0.0, self.keep_all_objs)
            if self.keep_all_objs:
                shp, rte = temp[0].reshape((1, -1)), temp[1].reshape((1, -1))
                new_k_rte = self.a_prime / self.b_prime + (shp / rte).sum(axis=1, keepdims=True)
            if update_existing:
                self.Theta[user_id] = Theta
                if self.keep_all_objs:
                    self.Gamma_shp[user_id] = shp
                    self.Gamma_rte[user_id] = rte
                    self.k_rte[user_id] = new_k_rte
            else:
                if self.reindex:
                    new_id = self.user_mapping_.shape[0]
                    self.user_mapping_ = np.r_[self.user_mapping_, np.array(user_id)]
                    if self.produce_dicts:
                        self.user_dict_[user_id] = new_id
                self.Theta = np.r_[self.Theta, Theta.reshape((1, self.k))]
                if self.keep_all_objs:
                    self.Gamma_shp = np.r_[self.Gamma_shp, shp]
                    self.Gamma_rte = np.r_[self.Gamma_rte, rte]
                    self.k_rte = np.r_[self.k_rte, new_k_rte.astype(self.k_rte.dtype)]
                self.nusers += 1
        self._drop_scorer()

        if self.keep_data:
            new_items = counts_df["ItemId"].to_numpy(copy=False)
            if update_existing:
                before = self._n_seen_by_user[user_id]
                st = self._st_ix_user[user_id]
                self.seen = np.r_[self.seen[:st], new_items, self.seen[st + before:]]
                self._n_seen_by_user[user_id] = counts_df.shape[0]
                self._st_ix_user[(user_id + 1):] += counts_df.shape[0] - before
            else:
                self._n_seen_by_user = np.r_[self._n_seen_by_user, np.array(counts_df.shape[0])]
                self._st_ix_user = np.r_[self._st_ix_user, self.seen.shape[0]]
                self.seen = np.r_[self.seen, new_items]
        return True

    # ------------------------------------------------------------------------------------------------
    def _row_ids(self, ids, mapping, lookup):
        """External ids (scalar or array) -> row numbers; unknown ids become -1."""
        if not np.isscalar(ids):
            ids = np.require(ids, requirements=["ENSUREARRAY"]).reshape(-1)
            assert ids.shape[0] > 0
            if not self.reindex:
                return ids
            if ids.shape[0] > 1:
                return _codes(ids, mapping)
            ids = ids[0]
        if self.reindex:
            if lookup is not None:
                ids = lookup.get(ids, -1)
            else:
                ids = _codes(np.array([ids]), mapping)[0]
        return np.array([ids])

    def predict(self, user, item):
        """Expected count(s) Theta[user] . Beta[item] for one pair or for aligned arrays of pairs;
        NaN where the user or item was not in the training data."""
        assert self.is_fitted
        user = self._row_ids(user, self.user_mapping_, self.user_dict_)
        item = self._row_ids(item, self.item_mapping_, self.item_dict_)
        assert user.shape[0] == item.shape[0]

        if user.shape[0] == 1:
            if user[0] == -1 or item[0] == -1:
                return np.nan
            return self.Theta[user].dot(self.Beta[item].T).reshape(-1)[0]

        sc = self._scorer()
        unknown = (user == -1) | (item == -1)
        if unknown.sum() == 0:
            return sc.predict(np.asarray(user, dtype=np.int64), np.asarray(item, dtype=np.int64))
        out = np.full(user.shape[0], np.nan, dtype=self.Theta.dtype)
        if (~unknown).sum() > 0:
            out[~unknown] = sc.predict(np.asarray(user[~unknown], dtype=np.int64), np.asarray(item[~unknown], dtype=np.int64))
        return out

    def topN(self, user, n=10, exclude_seen=True, items_pool=None):
        """Top-n items for a user by predicted count, optionally excluding her training items and/or
        restricted to `items_pool` (reference hpfrec/__init__.py:1296-1396)."""
        if isinstance(n, float):
            n = int(n)
        assert isinstance(n, int)
        if self.reindex:
            if self.produce_dicts:
                try:
                    user = self.user_dict_[user]
                except Exception:
                    raise ValueError("Can only predict for users who were in the training set.")
            else:
                user = _codes(np.array([user]), self.user_mapping_)[0]
                if user == -1:
                    raise ValueError("Can only predict for users who were in the training set.")
        if exclude_seen and not self.keep_data:
            raise Exception("Can only exclude seen items when passing 'keep_data=True' to .fit")

        def seen_by_user():
            st = int(self._st_ix_user[user])      # may be an unsigned 64-bit scalar after an SVI fit
            return self.seen[st: st + int(self._n_seen_by_user[user])]

        def back(rows):
            return self.item_mapping_[rows] if self.reindex else rows

        sc = self._scorer()
        seen = seen_by_user() if exclude_seen else None
        if items_pool is None:
            # all items, best first; the user's training items are masked on the device (the reference takes the
            # n + n_seen best with argpartition and removes the seen ones with setdiff1d: the same set)
            return back(sc.topn(user, min(n, self.Beta.shape[0]), seen=seen))

        items_pool = np.require(items_pool, requirements=["ENSUREARRAY"]).reshape(-1)
        pool_rows = items_pool
        if self.reindex:
            pool_rows = _codes(items_pool, self.item_mapping_)
            missing = pool_rows == -1
            if missing.sum() > 0:
                pool_rows = pool_rows[~missing]
                warnings.warn("There were %d entries from 'item_pool'"
                              "that were not in the training data and will be exluded." % int(missing.sum()))
            if pool_rows.shape[0] == 0:
                raise ValueError("No items to recommend.")
            if pool_rows.shape[0] == 1:
                raise ValueError("Only 1 item to recommend.")
        n = int(np.min([n, items_pool.shape[0]]))
        if exclude_seen:
            pool_rows = np.unique(pool_rows)   # the reference's setdiff1d de-duplicates the pool on this path
        return back(sc.topn(user, n, pool=pool_rows, seen=seen))

    def eval_llk(self, input_df, full_llk=False):
        """Poisson log-likelihood (plus constant unless full_llk) of the given triplets restricted to
        users/items known to the model: {'llk': value, 'nobs': rows used}
        (reference hpfrec/__init__.py:1399-1446)."""
        assert self.is_fitted
        self._process_valset(input_df, valset=False)
        vs = self.val_set
        o = self._scorer().llk(vs["UserId"].to_numpy(copy=False).astype(np.int64), vs["ItemId"].to_numpy(copy=False).astype(np.int64),
                               vs["Count"].to_numpy(copy=False), full_llk)
        out = {'llk': np.longdouble(o[0]) - np.longdouble(o[2]),   # calc_llk, pxi:525-534
               'nobs': vs.shape[0]}
        del self.val_set
        return out

    # ------------------------------------------------------------------------------------------------
    def _print_st_msg(self):
        print("**********************************")
        print("Hierarchical Poisson Factorization")
        print("**********************************")
        print("")

    def _print_data_info(self):
        print("Number of users: %d" % self.nusers)
        print("Number of items: %d" % self.nitems)
        print("Latent factors to use: %d" % self.k)
        print("")


def trim_cache():
    """Returns the device blocks the engine's allocator keeps cached between fits to the CUDA driver."""
    from . import _lib
    _lib.check(_lib.load().hpf_trim_cache())

