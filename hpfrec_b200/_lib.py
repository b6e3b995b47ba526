"""ctypes binding of the C ABI in include/hpf_b200.h (hpfrec_b200/_lib/libhpf_b200.so).

There is NO fallback: if the shared library is missing or a CUDA call fails, the caller gets an
exception.  Nothing here imports the CPU oracle.
"""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "_lib", "libhpf_b200.so")

# every symbol include/hpf_b200.h declares, with its argument types
_c = ctypes
_P = _c.c_void_p
_I32 = _c.c_int32
_I64 = _c.c_int64
_D = _c.c_double
SIGNATURES = {
    "hpf_abi_version": ([], _c.c_int),
    "hpf_last_error": ([], _c.c_char_p),
    "hpf_create": ([_c.POINTER(_P), _I64, _I64, _I32, _I32, _I32], _c.c_int),
    "hpf_destroy": ([_P], _c.c_int),
    "hpf_set_hyper": ([_P, _D, _D, _D, _D, _D, _D], _c.c_int),
    "hpf_set_constants": ([_P, _D, _D, _D, _D, _D, _D], _c.c_int),
    "hpf_set_stream": ([_P, _P], _c.c_int),
    "hpf_set_option": ([_P, _c.c_char_p, _D], _c.c_int),
    "hpf_load_state": ([_P, _P, _P, _P, _P, _P, _P], _c.c_int),
    "hpf_export_state": ([_P, _P, _P, _P, _P, _P, _P, _P, _P], _c.c_int),
    "hpf_load_coo": ([_P, _P, _P, _P, _I64, _I32], _c.c_int),
    "hpf_step_full": ([_P, _I32], _c.c_int),
    "hpf_sweep": ([_P], _c.c_int),
    "hpf_sweep_side": ([_P, _I32], _c.c_int),
    "hpf_update_users": ([_P], _c.c_int),
    "hpf_update_users_ex": ([_P, _c.c_int32], _c.c_int),
    "hpf_update_items": ([_P], _c.c_int),
    "hpf_item_pass_with_user_update": ([_P, _I32], _c.c_int),
    "hpf_partials": ([_P, _c.POINTER(_P), _c.POINTER(_I64), _c.POINTER(_P), _c.POINTER(_I64)], _c.c_int),
    "hpf_peer_export": ([_P, _P], _c.c_int),
    "hpf_peer_attach": ([_P, _I32, _I32, _P], _c.c_int),
    "hpf_item_buffer_bytes": ([_P, _c.POINTER(_I64)], _c.c_int),
    "hpf_adopt_item_buffers": ([_P, _c.POINTER(_P)], _c.c_int),
    "hpf_peer_attach_ptrs": ([_P, _I32, _I32, _c.POINTER(_P), _c.POINTER(_P)], _c.c_int),
    "hpf_update_items_peer": ([_P, _I32], _c.c_int),
    "hpf_reduce_items_peer": ([_P, _P], _c.c_int),
    "hpf_peer_finish": ([_P], _c.c_int),
    "hpf_beta_colsum": ([_P, _c.POINTER(_P), _c.POINTER(_I64)], _c.c_int),
    "hpf_step_batch": ([_P, _P, _P, _P, _I64, _P, _I64, _P, _I64, _I32, _I32, _D, _D, _I32], _c.c_int),
    "hpf_step_batch_ids": ([_P, _P, _I64, _I32, _I32, _D, _D, _I32], _c.c_int),
    "hpf_step_epoch_ids": ([_P, _P, _I64, _I32, _I64, _I32, _D], _c.c_int),
    "hpf_llk": ([_P, _P, _P, _P, _I64, _I32, _I32, _c.POINTER(_D)], _c.c_int),
    "hpf_llk_train": ([_P, _I32, _c.POINTER(_D)], _c.c_int),
    "hpf_predict": ([_P, _P, _P, _I64, _I32, _P], _c.c_int),
    "hpf_scorer_create": ([_c.POINTER(_P), _P, _P, _I64, _I64, _I32, _I32, _I32], _c.c_int),
    "hpf_scorer_destroy": ([_P], _c.c_int),
    "hpf_scorer_predict": ([_P, _P, _P, _I64, _I32, _P], _c.c_int),
    "hpf_scorer_llk": ([_P, _P, _P, _P, _I64, _I32, _I32, _c.POINTER(_D)], _c.c_int),
    "hpf_scorer_topn": ([_P, _I64, _I32, _P, _I64, _P, _I64, _I32, _c.POINTER(_I64), _P, _c.POINTER(_I32)], _c.c_int),
    "hpf_update_shapes": ([_I32, _I32, _I32, _P, _P, _P, _P, _P, _P, _P, _P, _I64, _I64, _I64, _I32, _D, _D],
                          _c.c_int),
    "hpf_digamma": ([_I32, _I32, _P, _P, _I64], _c.c_int),
    "hpf_factorize": ([_I32, _P, _I64, _I32, _P, _I32, _P, _c.POINTER(_I64)], _c.c_int),
    "hpf_csr_metadata": ([_I32, _P, _P, _I64, _I32, _I64, _I64, _P, _P, _I32, _c.POINTER(_I64)], _c.c_int),
    "hpf_trim_cache": ([], _c.c_int),
    "hpf_launch_count": ([_P, _c.POINTER(_I64)], _c.c_int),
    "hpf_phase_ms": ([_P, _c.POINTER(_D), _c.POINTER(_I64)], _c.c_int),
    "hpf_ld": ([_P, _c.POINTER(_I32)], _c.c_int),
    "hpf_describe": ([_P, _c.c_char_p, _I64], _c.c_int),
}

_lib = None


class HPFError(RuntimeError):
    """A call into libhpf_b200.so returned a non-zero status."""

    def __init__(self, code, msg):
        super().__init__("libhpf_b200 error %d: %s" % (code, msg))
        self.code = code


def load():
    """Loads (once) and returns the ctypes library.  Raises if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            "hpfrec_b200: %s not found. Build it with `python -m hpfrec_b200.build` (needs nvcc). "
            "This package has no CPU fallback." % LIB_PATH)
    lib = ctypes.CDLL(LIB_PATH)
    for name, (argtypes, restype) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the header and the library disagree
        fn.argtypes = argtypes
        fn.restype = restype
    _lib = lib
    return lib


def check(code):
    if code != 0:
        raise HPFError(code, load().hpf_last_error().decode("utf-8", "replace"))
