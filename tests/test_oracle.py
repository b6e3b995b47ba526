"""CPU tests (-m "not gpu"): the numpy oracle against (1) the committed golden vectors generated from
the compiled reference and (2) the compiled reference itself when oracle/_ref is present."""
import numpy as np
import pytest

from conftest import STATE_KEYS, relerr
from oracle import hpf_oracle as O
from oracle import ref_loader as R


@pytest.mark.parametrize("its,tol", [(1, 1e-13), (2, 1e-13), (10, 1e-12), (100, 1e-9)])
def test_full_batch_vs_golden(golden_full, its, tol):
    g = golden_full
    st = O.fit_full(g["Y"], g["ix_u"], g["ix_i"], 100, 100, 10, its, seed=123)
    for key in STATE_KEYS:
        assert relerr(st[key], g["it%d_%s" % (its, key)]) < tol, key
    assert int(g["it%d_niter" % its]) == its - 1     # reference returns the last loop index (pxi:418)


def test_survey_known_answer(golden_full):
    # SURVEY.md §8c "survey-session golden for the toy": regenerated, not trusted blindly
    g = golden_full
    np.testing.assert_allclose(g["it100_Theta"][0, :4], [0.34964709, 0.49778945, 0.13479044, 0.11405867], rtol=0, atol=1e-8)
    np.testing.assert_allclose(g["it100_Beta"][0, :4], [0.5729193, 0.73805154, 0.01979324, 0.9237548], rtol=0, atol=1e-7)
    assert abs(float(g["llk_full"]) - (-9120.387021600267)) < 1e-6
    assert g["Y"].shape[0] == 6347


def test_odd_shapes_vs_golden(golden_odd):
    g = golden_odd
    for its, tol in ((1, 1e-13), (25, 1e-11)):
        st = O.fit_full(g["Y"], g["ix_u"], g["ix_i"], int(g["nU"]), int(g["nI"]), int(g["k"]), its, seed=5,
                        a=0.5, a_prime=0.4, b_prime=1.3, c=0.6, c_prime=0.2, d_prime=0.8)
        for key in STATE_KEYS:
            assert relerr(st[key], g["it%d_%s" % (its, key)]) < tol, key


@pytest.mark.parametrize("name,upb,ipb", [("users", 20, 0), ("items", 0, 30), ("both", 20, 30)])
def test_svi_vs_golden(golden_svi, name, upb, ipb):
    g = golden_svi
    st = O.fit_svi(g["Y"], g["ix_u"], g["ix_i"], 100, 100, 10, 8, upb, ipb, seed=123)
    for key in STATE_KEYS:
        assert relerr(st[key], g["%s_%s" % (name, key)]) < 1e-12, key


def test_partial_fit_vs_golden(golden_full, golden_pf):
    g, p = golden_full, golden_pf
    u, i, y = g["ix_u"], g["ix_i"], g["Y"]
    st = O.initialize_parameters(100, 100, 10, 123, 0.3, 1.0, 0.3, 1.0)
    for call in range(5):
        kind = str(p["call%d_kind" % call])
        ids = p["call%d_ids" % call]
        sel = np.isin(u, ids) if kind == "users" else np.isin(i, ids)
        ub, ib, yb = u[sel], i[sel], y[sel]
        users, items = np.unique(ub), np.unique(ib)
        O.partial_fit_step(st, yb, ub, ib, users, items, kind == "users", float(p["call%d_rho" % call]),
                           100.0 / users.shape[0])
        for key in STATE_KEYS:
            assert relerr(st[key], p["call%d_%s" % (call, key)]) < 1e-12, (call, key)


def test_scores_vs_golden(golden_full):
    g = golden_full
    u, i, y = g["ix_u"], g["ix_i"], g["Y"]
    T, B = g["it100_Theta"], g["it100_Beta"]
    assert abs(float(O.calc_llk(y, u, i, T, B, True)) - float(g["llk_full"])) < 1e-8
    assert abs(float(O.calc_llk(y, u, i, T, B, False)) - float(g["llk_part"])) < 1e-8
    assert relerr(O.predict_arr(T, B, u, i), g["pred"]) < 1e-13


@pytest.mark.skipif(not R.available(), reason="oracle/_ref not built (needs /root/reference)")
def test_oracle_vs_compiled_reference_random():
    """Fresh seeded inputs straight through the compiled reference (not only the stored vectors)."""
    mod = R.load(False)
    u, i, y = O.synth_coo(300, 200, 5000, seed=3)
    for its in (1, 5):
        r = R.ref_fit_hpf(mod, y, u, i, 300, 200, 12, its, seed=77)
        st = O.fit_full(y, u, i, 300, 200, 12, its, seed=77)
        for key in STATE_KEYS:
            assert relerr(st[key], r[key]) < 1e-11, key


@pytest.mark.skipif(not R.available(), reason="oracle/_ref not built (needs /root/reference)")
def test_reference_has_openmp():
    # the trap of SURVEY.md §8c: a build without libgomp silently runs single-threaded
    import subprocess, glob, os
    so = glob.glob(os.path.join(os.path.dirname(R.__file__), "_ref", "hpfrec", "cython_loops_double*.so"))[0]
    out = subprocess.run(["ldd", so], stdout=subprocess.PIPE, text=True).stdout
    assert "libgomp" in out


def test_synth_generator_shape():
    u, i, y = O.synth_coo(2000, 700, 30000, seed=1)
    assert u.shape == i.shape == y.shape == (30000,)
    assert u.max() < 2000 and i.max() < 700 and y.min() >= 1
    assert np.unique(u * 700 + i).shape[0] == 30000          # de-duplicated
