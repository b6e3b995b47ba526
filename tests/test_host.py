"""CPU tests (-m "not gpu"): the C-ABI library loads and exports every declared symbol (no compute
calls), host-side logic of the Python mirror, and that the product never falls back to a CPU path."""
import ctypes
import os
import re

import numpy as np
import pandas as pd
import pytest

from conftest import ROOT, relerr
from oracle import hpf_oracle as O


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "hpf_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(hpf_[a-z_0-9]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    from hpfrec_b200 import _lib
    lib = ctypes.CDLL(_lib.LIB_PATH)
    names = _declared_symbols()
    assert len(names) >= 20
    for name in names:
        assert hasattr(lib, name), "libhpf_b200.so lacks %s" % name
    # and the ctypes table binds exactly the header's functions
    assert sorted(_lib.SIGNATURES) == names
    assert _lib.load().hpf_abi_version() == 1


def test_no_cpu_fallback_without_gpu():
    """Without a CUDA device every compute entry point must fail loudly, not compute on the host."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from hpfrec_b200 import _lib
    from hpfrec_b200.engine import Engine
    with pytest.raises(_lib.HPFError):
        Engine(10, 10, 4)
    from hpfrec_b200 import HPF
    df = O.readme_toy()
    with pytest.raises(_lib.HPFError):
        HPF(k=5, verbose=False, maxiter=2, check_every=None).fit(df)


def test_product_does_not_import_oracle():
    pkg = os.path.join(ROOT, "hpfrec_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".inl")):
                src = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in src and "from oracle" not in src, f
                assert "hpf_oracle" not in src and "ref_loader" not in src, f


def test_initialize_parameters_matches_reference_stream(golden_full):
    from hpfrec_b200.loops import cuda_loops_double, cuda_loops_float
    Theta, Beta = np.empty((100, 10)), np.empty((100, 10))
    got = cuda_loops_double.initialize_parameters(Theta, Beta, 123, 0.3, 0.3, 1.0, 0.3, 0.3, 1.0)
    ref = O.initialize_parameters(100, 100, 10, 123, 0.3, 1.0, 0.3, 1.0)
    for arr, key in zip(got, ("Gamma_shp", "Gamma_rte", "Lambda_shp", "Lambda_rte", "k_rte", "t_rte")):
        assert np.array_equal(arr, ref[key]), key
    assert np.array_equal(Theta, ref["Theta"])
    # float instantiation draws float32 directly (different bit-stream)
    Tf, Bf = np.empty((7, 3), np.float32), np.empty((5, 3), np.float32)
    gf = cuda_loops_float.initialize_parameters(Tf, Bf, 9, 0.3, 0.3, 1.0, 0.3, 0.3, 1.0)
    rf = O.initialize_parameters(7, 5, 3, 9, 0.3, 1.0, 0.3, 1.0, dtype=np.float32)
    assert gf[0].dtype == np.float32 and np.array_equal(gf[0], rf["Gamma_shp"])
    assert np.array_equal(gf[3], rf["Lambda_rte"])


def test_rows_in_order():
    from hpfrec_b200.loops import _rows_in_order
    indptr = np.array([0, 3, 3, 7, 8], dtype=np.int64)
    pos, cnt = _rows_in_order(indptr, np.array([2, 0, 1, 3]))
    assert pos.tolist() == [3, 4, 5, 6, 0, 1, 2, 7]
    assert cnt.tolist() == [4, 3, 0, 1]
    pos, cnt = _rows_in_order(indptr, np.array([1]))
    assert pos.shape[0] == 0


def test_constructor_validation_matches_reference_rules():
    from hpfrec_b200 import HPF
    m = HPF()
    assert (m.k, m.a, m.maxiter, m.check_every, m.stop_crit, m.use_float) == (30, 0.3, 100, 10, "maxiter", True)
    assert HPF(verbose=False).check_every == 0                 # reference __init__.py:291-292
    assert HPF(reindex=False).produce_dicts is False           # reference __init__.py:346-347
    assert HPF(users_per_batch=20.0).users_per_batch == 20
    assert HPF(a=1).a == 1.0
    with pytest.raises(AssertionError):
        HPF(k=0)
    with pytest.raises(AssertionError):
        HPF(a=-1.0)
    with pytest.raises(AssertionError):
        HPF(stop_crit="nope")
    with pytest.raises(AssertionError):
        HPF(check_every=200, maxiter=100)                       # reference __init__.py:274
    with pytest.raises(ValueError):
        HPF(maxiter=None)
    with pytest.raises(ValueError):
        HPF(stop_crit="train-llk", check_every=None)
    with pytest.raises(ValueError):
        HPF(step_size=0.5)
    with pytest.raises(ValueError):
        HPF(stop_crit="val-llk").fit(O.readme_toy())


def test_process_data_reindex_and_filtering():
    """Host logic of HPF._process_data that needs no device: string ids stay with pandas, zero counts are dropped,
    dtypes follow use_float, a COO input switches reindexing off."""
    from hpfrec_b200 import HPF
    df = pd.DataFrame({"UserId": ["b", "a", "b", "c"], "ItemId": ["x", "x", "z", "y"], "Count": [1, 2, 0, 3]})
    m = HPF(k=3, verbose=False)
    with pytest.warns(UserWarning):
        m._process_data(df)
    assert m.nusers == 3 and m.nitems == 2                    # the zero-count row (b,z) is dropped
    assert m.user_mapping_.tolist() == ["b", "a", "c"] and m.item_mapping_.tolist() == ["x", "y"]
    assert m.input_df["UserId"].tolist() == [0, 1, 2]
    assert m.input_df["Count"].dtype == np.float32

    from scipy.sparse import coo_array
    X = coo_array((np.array([1., 2.]), (np.array([0, 4]), np.array([1, 2]))), shape=(6, 5))
    m2 = HPF(k=3, verbose=False, use_float=False)
    m2._process_data(X)
    assert m2.reindex is False and (m2.nusers, m2.nitems) == (6, 5)
    assert m2.input_df["Count"].dtype == np.float64


def test_integer_id_ingest_needs_the_device():
    """Integer id columns are factorized by the CUDA library; without a GPU that fails loudly (no pandas fallback)."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    from hpfrec_b200 import HPF
    from hpfrec_b200._lib import HPFError
    df = pd.DataFrame({"UserId": [3, 1, 3], "ItemId": [10, 10, 30], "Count": [1, 2, 3]})
    with pytest.raises(HPFError):
        HPF(k=3, verbose=False)._process_data(df)


def test_oracle_llk_shortcut_is_consistent():
    rng = np.random.default_rng(0)
    T, B = rng.random((6, 3)), rng.random((5, 3))
    u, i = np.array([0, 1, 5, 2]), np.array([4, 0, 0, 3])
    y = np.array([1., 2., 1., 4.])
    l, rmse = O.train_llk(T, B, y, u, i)
    full = (T @ B.T).sum()
    l2, _ = O.llk_plus_rmse(T, B, y, u, i)
    assert abs(float(l2 - l) - full) < 1e-10
