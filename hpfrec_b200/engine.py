"""Thin object wrapper over the C ABI (include/hpf_b200.h).  Arrays may be C-contiguous numpy arrays
(host) or torch CUDA tensors (device); only their raw pointers cross into the library."""
import ctypes

import numpy as np

from . import _lib

REAL_DTYPES = {4: np.float32, 8: np.float64}


def _ptr(x, dtype=None, name="array"):
    """Raw pointer of a numpy array / torch tensor; None -> NULL.  Keeps no reference."""
    if x is None:
        return None
    if isinstance(x, np.ndarray):
        if not x.flags["C_CONTIGUOUS"]:
            raise ValueError("%s must be C-contiguous" % name)
        if dtype is not None and x.dtype != np.dtype(dtype):
            raise ValueError("%s must have dtype %s, got %s" % (name, np.dtype(dtype), x.dtype))
        return ctypes.c_void_p(x.ctypes.data)
    if hasattr(x, "data_ptr"):  # torch tensor
        if not x.is_contiguous():
            raise ValueError("%s must be contiguous" % name)
        if dtype is not None:
            import torch
            want = {np.dtype(np.float32): torch.float32, np.dtype(np.float64): torch.float64,
                    np.dtype(np.int32): torch.int32, np.dtype(np.int64): torch.int64}[np.dtype(dtype)]
            if x.dtype != want:
                raise ValueError("%s must have dtype %s, got %s" % (name, want, x.dtype))
        return ctypes.c_void_p(x.data_ptr())
    raise TypeError("%s: expected numpy array or torch tensor" % name)


def _index_bytes(x):
    if isinstance(x, np.ndarray):
        sz = x.dtype.itemsize
        if x.dtype.kind not in "iu" or sz not in (4, 8):
            raise ValueError("index arrays must be 32- or 64-bit integers, got %s" % x.dtype)
        return sz
    return x.element_size()


def as_index(x):
    """Host index array -> contiguous int32/int64/uint64 numpy array the C ABI accepts as-is."""
    x = np.asarray(x)
    if x.dtype.kind not in "iu" or x.dtype.itemsize not in (4, 8):
        x = x.astype(np.int64)
    return np.ascontiguousarray(x).reshape(-1)


class Engine:
    """One device-resident HPF state (+ optionally the training triples)."""

    def __init__(self, nU, nI, k, real_bytes=4, device=0, stream=None):
        self._lib = _lib.load()
        self._h = ctypes.c_void_p()
        self.nU, self.nI, self.k, self.real_bytes = int(nU), int(nI), int(k), int(real_bytes)
        self.dtype = np.dtype(REAL_DTYPES[self.real_bytes])
        _lib.check(self._lib.hpf_create(ctypes.byref(self._h), self.nU, self.nI, self.k, self.real_bytes,
                                        int(device)))
        if stream is not None:
            self.set_stream(stream)

    # -- life cycle ---------------------------------------------------------------------------
    def close(self):
        if getattr(self, "_h", None) is not None and self._h.value:
            self._lib.hpf_destroy(self._h)
            self._h = ctypes.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    # -- configuration --------------------------------------------------------------------------
    def set_hyper(self, a, a_prime, b_prime, c, c_prime, d_prime):
        _lib.check(self._lib.hpf_set_hyper(self._h, a, a_prime, b_prime, c, c_prime, d_prime))

    def set_constants(self, a, c, k_shp, t_shp, add_k_rte, add_t_rte):
        _lib.check(self._lib.hpf_set_constants(self._h, a, c, k_shp, t_shp, add_k_rte, add_t_rte))

    def set_stream(self, stream):
        """`stream`: an int/ctypes pointer (cudaStream_t) or a torch.cuda.Stream."""
        ptr = getattr(stream, "cuda_stream", stream)
        _lib.check(self._lib.hpf_set_stream(self._h, ctypes.c_void_p(int(ptr) if ptr else 0)))

    def set_option(self, name, value):
        _lib.check(self._lib.hpf_set_option(self._h, name.encode(), float(value)))

    @property
    def ld(self):
        out = ctypes.c_int32()
        _lib.check(self._lib.hpf_ld(self._h, ctypes.byref(out)))
        return out.value

    def describe(self):
        """Resolved configuration as a dict of strings (row stride, sweep mode, kernel, shape, panels)."""
        buf = ctypes.create_string_buffer(512)
        _lib.check(self._lib.hpf_describe(self._h, buf, 512))
        return dict(item.split("=", 1) for item in buf.value.decode().split())

    @property
    def launch_count(self):
        out = ctypes.c_int64()
        _lib.check(self._lib.hpf_launch_count(self._h, ctypes.byref(out)))
        return out.value

    def phase_ms(self):
        """(ms[4], iterations) accumulated since set_option('timing', 1): item-major pass, user-major
        pass, user update, item update."""
        out = (ctypes.c_double * 4)()
        n = ctypes.c_int64()
        _lib.check(self._lib.hpf_phase_ms(self._h, out, ctypes.byref(n)))
        return list(out), n.value

    # -- state ------------------------------------------------------------------------------------
    def load_state(self, Gamma_shp, Gamma_rte, Lambda_shp, Lambda_rte, k_rte, t_rte):
        dt = self.dtype
        args = [_ptr(x, dt, n) for x, n in ((Gamma_shp, "Gamma_shp"), (Gamma_rte, "Gamma_rte"),
                                            (Lambda_shp, "Lambda_shp"), (Lambda_rte, "Lambda_rte"),
                                            (k_rte, "k_rte"), (t_rte, "t_rte"))]
        _lib.check(self._lib.hpf_load_state(self._h, *args))

    def export_state(self, Gamma_shp=None, Gamma_rte=None, Lambda_shp=None, Lambda_rte=None,
                     k_rte=None, t_rte=None, Theta=None, Beta=None):
        dt = self.dtype
        args = [_ptr(x, dt) for x in (Gamma_shp, Gamma_rte, Lambda_shp, Lambda_rte, k_rte, t_rte, Theta, Beta)]
        _lib.check(self._lib.hpf_export_state(self._h, *args))

    def export_all(self):
        """Convenience: fresh numpy arrays for the six state arrays + Theta, Beta."""
        dt, k = self.dtype, self.k
        out = dict(Gamma_shp=np.empty((self.nU, k), dt), Gamma_rte=np.empty((self.nU, k), dt),
                   Lambda_shp=np.empty((self.nI, k), dt), Lambda_rte=np.empty((self.nI, k), dt),
                   k_rte=np.empty((self.nU, 1), dt), t_rte=np.empty((self.nI, 1), dt),
                   Theta=np.empty((self.nU, k), dt), Beta=np.empty((self.nI, k), dt))
        self.export_state(**out)
        return out

    # -- data -------------------------------------------------------------------------------------
    def load_coo(self, ix_u, ix_i, Y):
        ib = _index_bytes(ix_u)
        if _index_bytes(ix_i) != ib:
            raise ValueError("ix_u and ix_i must have the same integer width")
        n = int(Y.shape[0])
        _lib.check(self._lib.hpf_load_coo(self._h, _ptr(ix_u, name="ix_u"), _ptr(ix_i, name="ix_i"),
                                          _ptr(Y, self.dtype, "Y"), n, ib))

    # -- full batch ---------------------------------------------------------------------------------
    def step_full(self, niter=1):
        _lib.check(self._lib.hpf_step_full(self._h, int(niter)))

    def sweep(self):
        _lib.check(self._lib.hpf_sweep(self._h))

    def sweep_side(self, side):
        """side 0: item-major pass (item-side partial sums), side 1: user-major pass."""
        _lib.check(self._lib.hpf_sweep_side(self._h, int(side)))

    def update_users(self, materialize=True):
        _lib.check(self._lib.hpf_update_users_ex(self._h, int(bool(materialize))))

    def item_pass_with_user_update(self, materialize=True):
        """Item-major pass with the user update running under it on a second stream (after sweep_side(1))."""
        _lib.check(self._lib.hpf_item_pass_with_user_update(self._h, int(bool(materialize))))

    def update_items(self):
        _lib.check(self._lib.hpf_update_items(self._h))

    def partials(self):
        """(item_sums_ptr, count, theta_colsum_ptr, count): raw device pointers for the all-reduce."""
        p1, n1, p2, n2 = ctypes.c_void_p(), ctypes.c_int64(), ctypes.c_void_p(), ctypes.c_int64()
        _lib.check(self._lib.hpf_partials(self._h, ctypes.byref(p1), ctypes.byref(n1), ctypes.byref(p2),
                                          ctypes.byref(n2)))
        return p1.value, n1.value, p2.value, n2.value

    # -- NVLink peer memory (one process per GPU) --------------------------------------------------
    PEER_HANDLE_BYTES = 64 * 5

    def peer_export(self):
        buf = ctypes.create_string_buffer(self.PEER_HANDLE_BYTES)
        _lib.check(self._lib.hpf_peer_export(self._h, ctypes.cast(buf, ctypes.c_void_p)))
        return bytes(buf.raw)

    def peer_attach(self, rank, world, all_handles):
        """all_handles: the concatenation, in rank order, of every rank's peer_export() bytes."""
        if len(all_handles) != world * self.PEER_HANDLE_BYTES:
            raise ValueError("expected %d bytes of IPC handles" % (world * self.PEER_HANDLE_BYTES))
        buf = ctypes.create_string_buffer(bytes(all_handles), len(all_handles))
        _lib.check(self._lib.hpf_peer_attach(self._h, int(rank), int(world), ctypes.cast(buf, ctypes.c_void_p)))

    N_PEER_BUFFERS = 5

    def item_buffer_bytes(self):
        """Sizes of the five item-side buffers (item sums, item factors, t_rte, Lambda_shp, Lambda_rte)."""
        out = (ctypes.c_int64 * self.N_PEER_BUFFERS)()
        _lib.check(self._lib.hpf_item_buffer_bytes(self._h, out))
        return list(out)

    def adopt_item_buffers(self, ptrs):
        """Use caller-owned device memory (e.g. symmetric memory) for the five item-side buffers."""
        arr = (ctypes.c_void_p * self.N_PEER_BUFFERS)(*[int(p) for p in ptrs])
        _lib.check(self._lib.hpf_adopt_item_buffers(self._h, arr))

    def peer_attach_ptrs(self, rank, world, peer_ptrs, mc_ptrs=None):
        """peer_ptrs[p][b]: rank p's buffer b as mapped in THIS process; mc_ptrs[b]: multicast mapping or None."""
        flat = (ctypes.c_void_p * (world * self.N_PEER_BUFFERS))(*[int(x) for row in peer_ptrs for x in row])
        mc = None
        if mc_ptrs is not None:
            mc = (ctypes.c_void_p * self.N_PEER_BUFFERS)(*[int(x) for x in mc_ptrs])
        _lib.check(self._lib.hpf_peer_attach_ptrs(self._h, int(rank), int(world), flat, mc))

    def update_items_peer(self, materialize=True):
        _lib.check(self._lib.hpf_update_items_peer(self._h, int(bool(materialize))))

    def reduce_items_peer(self, stream=None):
        """The reduce-scatter half of the exchange alone, on `stream` (raw cudaStream_t handle; None = the engine's)."""
        _lib.check(self._lib.hpf_reduce_items_peer(self._h, ctypes.c_void_p(int(stream) if stream else None)))

    def peer_finish(self):
        _lib.check(self._lib.hpf_peer_finish(self._h))

    def beta_colsum(self):
        p, n = ctypes.c_void_p(), ctypes.c_int64()
        _lib.check(self._lib.hpf_beta_colsum(self._h, ctypes.byref(p), ctypes.byref(n)))
        return p.value, n.value

    # -- minibatch ----------------------------------------------------------------------------------
    def step_batch(self, ix_u, ix_i, Y, users, items, user_batch, rho, mult, blend_all_rates):
        ib = _index_bytes(ix_u)
        for arr in (ix_i, users, items):
            if _index_bytes(arr) != ib:
                raise ValueError("all index arrays of a minibatch must have the same integer width")
        _lib.check(self._lib.hpf_step_batch(
            self._h, _ptr(ix_u, name="ix_u"), _ptr(ix_i, name="ix_i"), _ptr(Y, self.dtype, "Y"), int(Y.shape[0]),
            _ptr(users, name="users"), int(users.shape[0]), _ptr(items, name="items"), int(items.shape[0]),
            ib, int(bool(user_batch)), float(rho), float(mult), int(bool(blend_all_rates))))

    def step_batch_ids(self, ids, user_batch, rho, mult, blend_all_rates=False):
        """Minibatch update for the rows `ids` using the triples resident on the device."""
        _lib.check(self._lib.hpf_step_batch_ids(self._h, _ptr(ids, name="ids"), int(ids.shape[0]), _index_bytes(ids),
                                                int(bool(user_batch)), float(rho), float(mult),
                                                int(bool(blend_all_rates))))

    def step_epoch_ids(self, ids, batch_rows, user_batch, rho):
        """One SVI epoch: consecutive slices of `batch_rows` ids of the (shuffled) list are the minibatches.
        Asynchronous: returns with the epoch's kernels in flight on the engine's stream."""
        _lib.check(self._lib.hpf_step_epoch_ids(self._h, _ptr(ids, name="ids"), int(ids.shape[0]), _index_bytes(ids),
                                                int(batch_rows), int(bool(user_batch)), float(rho)))

    # -- metrics / scoring --------------------------------------------------------------------------
    def llk(self, ix_u, ix_i, Y, full_llk=False):
        out = (ctypes.c_double * 4)()
        _lib.check(self._lib.hpf_llk(self._h, _ptr(ix_u), _ptr(ix_i), _ptr(Y, self.dtype, "Y"), int(Y.shape[0]),
                                     _index_bytes(ix_u), int(bool(full_llk)), out))
        return list(out)

    def llk_train(self, full_llk=False):
        out = (ctypes.c_double * 4)()
        _lib.check(self._lib.hpf_llk_train(self._h, int(bool(full_llk)), out))
        return list(out)

    def predict(self, ix_u, ix_i, out=None):
        n = int(ix_u.shape[0])
        if out is None:
            out = np.empty(n, dtype=self.dtype)
        _lib.check(self._lib.hpf_predict(self._h, _ptr(ix_u), _ptr(ix_i), n, _index_bytes(ix_u),
                                         _ptr(out, self.dtype, "out")))
        return out


def update_shapes(G_sh, G_rt, L_sh, L_rt, Y, ix_u, ix_i, a, c, phi=None, device=0):
    """Stateless fused update_phi + update_G_n_L_sh (reference pxi:551 + pxi:613) on caller buffers:
    G_sh / L_sh are overwritten with a + sum(phi), c + sum(phi); optional phi (nY x k) is filled."""
    lib = _lib.load()
    dt = G_sh.dtype if isinstance(G_sh, np.ndarray) else None
    rb = G_sh.dtype.itemsize if isinstance(G_sh, np.ndarray) else G_sh.element_size()
    _lib.check(lib.hpf_update_shapes(rb, _index_bytes(ix_u), int(device), _ptr(G_sh, dt), _ptr(G_rt, dt),
                                     _ptr(L_sh, dt), _ptr(L_rt, dt), _ptr(phi, dt), _ptr(Y, dt), _ptr(ix_u),
                                     _ptr(ix_i), int(G_sh.shape[0]), int(L_sh.shape[0]), int(Y.shape[0]),
                                     int(G_sh.shape[1]), float(a), float(c)))


def digamma(x, device=0):
    """The engine's device digamma (for validation against scipy.special.psi)."""
    lib = _lib.load()
    x = np.ascontiguousarray(x)
    out = np.empty_like(x)
    _lib.check(lib.hpf_digamma(x.dtype.itemsize, int(device), _ptr(x), _ptr(out), int(x.size)))
    return out


def factorize(values, device=0):
    """pd.factorize of an integer id column on the device (hpf_factorize): returns (codes int64, uniques in the
    column's own dtype), codes numbered in order of first appearance exactly like pandas."""
    lib = _lib.load()
    v = np.ascontiguousarray(values).reshape(-1)
    if v.dtype.kind not in "iu" or v.dtype.itemsize not in (4, 8):
        raise ValueError("device factorize needs a 32- or 64-bit integer column, got %s" % v.dtype)
    n = int(v.shape[0])
    codes = np.empty(n, dtype=np.int64)
    uniques = np.empty(n, dtype=v.dtype)
    n_unique = ctypes.c_int64(0)
    _lib.check(lib.hpf_factorize(int(device), _ptr(v), n, v.dtype.itemsize, _ptr(codes), 8, _ptr(uniques),
                                 ctypes.byref(n_unique)))
    return codes, uniques[:n_unique.value].copy()


def csr_metadata(ix_u, ix_i, nU, nI, device=0):
    """(indptr, indices) of coo_array((.., (ix_u, ix_i)), shape=(nU, nI)).tocsr() built on the device
    (hpf_csr_metadata): duplicate pairs merged, item ids ascending within a user; int32 like scipy's at these sizes."""
    lib = _lib.load()
    u, i = as_index(ix_u), as_index(ix_i)
    if u.dtype.itemsize != i.dtype.itemsize:
        u, i = u.astype(np.int64), i.astype(np.int64)
    n = int(u.shape[0])
    indptr = np.empty(int(nU) + 1, dtype=np.int64)
    indices = np.empty(max(n, 1), dtype=np.int32)
    n_out = ctypes.c_int64(0)
    _lib.check(lib.hpf_csr_metadata(int(device), _ptr(u), _ptr(i), n, u.dtype.itemsize, int(nU), int(nI),
                                    _ptr(indptr), _ptr(indices), 4, ctypes.byref(n_out)))
    return indptr.astype(np.int32), indices[:n_out.value].copy()


class Scorer:
    """Theta and Beta resident on the device (hpf_scorer_*): predict / llk / topN of a fitted model without
    re-uploading the factors per call."""

    def __init__(self, Theta, Beta, device=0):
        self._lib = _lib.load()
        self._h = ctypes.c_void_p()
        Theta = np.ascontiguousarray(Theta)
        Beta = np.ascontiguousarray(Beta, dtype=Theta.dtype)
        if Theta.dtype not in (np.dtype(np.float32), np.dtype(np.float64)):
            raise TypeError("Theta/Beta must be float32 or float64")
        self.dtype = Theta.dtype
        self.nU, self.k = Theta.shape
        self.nI = Beta.shape[0]
        _lib.check(self._lib.hpf_scorer_create(ctypes.byref(self._h), _ptr(Theta), _ptr(Beta), self.nU, self.nI,
                                               self.k, self.dtype.itemsize, int(device)))

    def close(self):
        if getattr(self, "_h", None) is not None and self._h.value:
            self._lib.hpf_scorer_destroy(self._h)
            self._h = ctypes.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def predict(self, ix_u, ix_i):
        ix_u, ix_i = as_index(ix_u), as_index(ix_i)
        if ix_i.dtype != ix_u.dtype:
            ix_i = ix_i.astype(ix_u.dtype)
        out = np.empty(ix_u.shape[0], dtype=self.dtype)
        _lib.check(self._lib.hpf_scorer_predict(self._h, _ptr(ix_u), _ptr(ix_i), int(ix_u.shape[0]), _index_bytes(ix_u),
                                                _ptr(out)))
        return out

    def llk(self, ix_u, ix_i, Y, full_llk=False):
        """[sum Y log yhat (- lgamma(Y+1) if full_llk), sum (Y - yhat)^2, sum yhat]"""
        ix_u, ix_i = as_index(ix_u), as_index(ix_i)
        if ix_i.dtype != ix_u.dtype:
            ix_i = ix_i.astype(ix_u.dtype)
        Y = np.ascontiguousarray(Y, dtype=self.dtype)
        out = (ctypes.c_double * 3)()
        _lib.check(self._lib.hpf_scorer_llk(self._h, _ptr(ix_u), _ptr(ix_i), _ptr(Y), int(Y.shape[0]), _index_bytes(ix_u),
                                            int(bool(full_llk)), out))
        return list(out)

    def topn(self, user, n, pool=None, seen=None, with_scores=False):
        """Item rows of the n best items for user row `user`, best first (see hpf_scorer_topn)."""
        n = int(n)
        pool = None if pool is None else as_index(pool).astype(np.int64)
        seen = None if seen is None or len(seen) == 0 else as_index(seen).astype(np.int64)
        ids = np.empty(max(n, 1), dtype=np.int64)
        scores = np.empty(max(n, 1), dtype=self.dtype)
        n_out = ctypes.c_int32()
        _lib.check(self._lib.hpf_scorer_topn(
            self._h, int(user), n, _ptr(pool), 0 if pool is None else int(pool.shape[0]), _ptr(seen),
            0 if seen is None else int(seen.shape[0]), 8, ids.ctypes.data_as(ctypes.POINTER(ctypes.c_int64)), _ptr(scores),
            ctypes.byref(n_out)))
        ids, scores = ids[:n_out.value], scores[:n_out.value]
        return (ids, scores) if with_scores else ids

