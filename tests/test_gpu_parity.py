"""GPU parity tests (-m gpu): the CUDA path, called through the C ABI (ctypes), against the numpy
oracle on the same seeded inputs, against the committed golden vectors of the compiled reference,
and -- when oracle/_ref travelled to the box -- against the compiled reference itself.

Tolerances (BASELINE.json north_star: fitted Theta/Beta within 1e-5 relative of the reference's
fp64 deterministic run):
  * fp64 engine vs reference fp64: <= 1e-9 after 100 iterations (observed ~1e-12; summation order only)
  * fp32 engine: one sweep from identical state <= 1e-5 (SURVEY.md §7 "parity budget vs fp32")
"""
import numpy as np
import pytest

from conftest import STATE_KEYS, relerr
from oracle import hpf_oracle as O
from oracle import ref_loader as R

pytestmark = pytest.mark.gpu

HYP = dict(a=0.3, a_prime=0.3, b_prime=1.0, c=0.3, c_prime=0.3, d_prime=1.0)


def _engine_from(st, k, dtype, **opts):
    from hpfrec_b200.engine import Engine
    nU, nI = st["Gamma_shp"].shape[0], st["Lambda_shp"].shape[0]
    eng = Engine(nU, nI, k, np.dtype(dtype).itemsize)
    for name, val in opts.items():
        eng.set_option(name, val)
    c = lambda x: np.ascontiguousarray(x, dtype=dtype)
    eng.load_state(c(st["Gamma_shp"]), c(st["Gamma_rte"]), c(st["Lambda_shp"]), c(st["Lambda_rte"]),
                   c(st["k_rte"]), c(st["t_rte"]))
    return eng


def _fit_gpu(Y, u, i, nU, nI, k, its, seed, dtype=np.float64, hyp=HYP, **opts):
    st = O.initialize_parameters(nU, nI, k, seed, hyp["a_prime"], hyp["b_prime"], hyp["c_prime"], hyp["d_prime"], dtype)
    eng = _engine_from(st, k, dtype, **opts)
    eng.set_hyper(hyp["a"], hyp["a_prime"], hyp["b_prime"], hyp["c"], hyp["c_prime"], hyp["d_prime"])
    eng.load_coo(np.ascontiguousarray(u, np.int64), np.ascontiguousarray(i, np.int64), np.ascontiguousarray(Y, dtype))
    eng.step_full(its)
    out = eng.export_all()
    eng.close()
    return out


# ---------------------------------------------------------------------------------------------------
def test_digamma_vs_scipy():
    from scipy.special import psi
    from hpfrec_b200.engine import digamma
    x = np.concatenate([np.geomspace(1e-3, 1e7, 200001), np.linspace(0.25, 12, 100001),
                        1.4616321449683623 + np.linspace(-1e-3, 1e-3, 1001)])
    got = digamma(x)
    ref = psi(x)
    # abs err <= 1e-13 in fp64 (relative where |psi| is large)
    assert np.max(np.abs(got - ref) / np.maximum(1.0, np.abs(ref))) < 1e-13
    xf = x[x >= 0.01].astype(np.float32)
    gotf = digamma(xf)
    reff = psi(xf.astype(np.float64))
    assert np.max(np.abs(gotf - reff) / np.maximum(1.0, np.abs(reff))) < 2e-6


@pytest.mark.parametrize("its,tol", [(1, 1e-12), (2, 1e-12), (10, 1e-11), (100, 1e-9)])
@pytest.mark.parametrize("sweep", [0, 1])
def test_full_batch_fp64_vs_golden_reference(golden_full, its, tol, sweep):
    """Config C1 (README toy, k=10, fp64): all eight arrays vs the compiled reference's output."""
    g = golden_full
    out = _fit_gpu(g["Y"], g["ix_u"], g["ix_i"], 100, 100, 10, its, 123, sweep=sweep)
    for key in STATE_KEYS:
        assert relerr(out[key], g["it%d_%s" % (its, key)]) < tol, key


def test_full_batch_fp64_odd_shapes(golden_odd):
    """k=7 (not a multiple of the 16-byte pack), users/items without data, non-default hyper-parameters."""
    g = golden_odd
    hyp = dict(a=0.5, a_prime=0.4, b_prime=1.3, c=0.6, c_prime=0.2, d_prime=0.8)
    for its, tol in ((1, 1e-12), (25, 1e-10)):
        out = _fit_gpu(g["Y"], g["ix_u"], g["ix_i"], int(g["nU"]), int(g["nI"]), int(g["k"]), its, 5, hyp=hyp,
                       chunk=32)
        for key in STATE_KEYS:
            assert relerr(out[key], g["it%d_%s" % (its, key)]) < tol, key


def test_fp32_single_sweep_and_short_run(golden_full):
    g = golden_full
    y32 = g["Y"].astype(np.float32)
    one = _fit_gpu(y32, g["ix_u"], g["ix_i"], 100, 100, 10, 1, 123, dtype=np.float32)
    # one sweep from the identical (float32-drawn) start: vs the reference's float build and vs the
    # fp64 oracle started from the same float32 numbers
    assert relerr(one["Theta"], g["f32_it1_Theta"]) < 1e-5
    assert relerr(one["Beta"], g["f32_it1_Beta"]) < 1e-5
    st = O.initialize_parameters(100, 100, 10, 123, 0.3, 1.0, 0.3, 1.0, np.float32)
    st = {k_: v.astype(np.float64) for k_, v in st.items()}
    O.cavi_full_iteration(st, g["Y"], g["ix_u"], g["ix_i"], **HYP)
    assert relerr(one["Theta"], st["Theta"]) < 1e-5
    assert relerr(one["Beta"], st["Beta"]) < 1e-5
    ten = _fit_gpu(y32, g["ix_u"], g["ix_i"], 100, 100, 10, 10, 123, dtype=np.float32)
    assert relerr(ten["Theta"], g["f32_it10_Theta"]) < 2e-4      # fp32 noise amplification


@pytest.mark.parametrize("k", [4, 30, 50, 64, 100, 128])
@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_sweep_all_row_shapes_vs_oracle(k, dtype):
    """Every lane-group configuration (k -> padded row length) on ragged synthetic data, both sweep
    implementations, 2 iterations, vs the fp64 oracle from the same start."""
    nU, nI, nnz = 700, 300, 20000
    u, i, y = O.synth_coo(nU, nI, nnz, seed=k)
    st0 = O.initialize_parameters(nU, nI, k, 11, 0.3, 1.0, 0.3, 1.0, dtype)
    ref = {k_: v.astype(np.float64) for k_, v in st0.items()}
    for _ in range(2):
        O.cavi_full_iteration(ref, y, u, i, **HYP)
    tol = 1e-11 if dtype == np.float64 else 2e-5
    for sweep in (0, 1):
        out = _fit_gpu(y, u, i, nU, nI, k, 2, 11, dtype=dtype, sweep=sweep, panel_mb=0.05, chunk=64)
        for key in STATE_KEYS:
            assert relerr(out[key], ref[key]) < tol, (key, sweep)


def test_update_shapes_one_shot_matches_reference_loops(golden_full):
    """hpf_update_shapes == update_phi (pxi:551) + update_G_n_L_sh (pxi:613) on host buffers with the
    reference's 8-byte indices, including the materialised phi."""
    from hpfrec_b200.engine import update_shapes
    g = golden_full
    u, i, y = g["ix_u"].astype(np.uint64), g["ix_i"].astype(np.uint64), g["Y"]
    st = O.initialize_parameters(100, 100, 10, 123, 0.3, 1.0, 0.3, 1.0)
    phi_ref = O.phi_rows(st["Gamma_shp"], st["Gamma_rte"], st["Lambda_shp"], st["Lambda_rte"], y,
                         g["ix_u"], g["ix_i"], True)
    G_ref, L_ref = np.full((100, 10), 0.3), np.full((100, 10), 0.3)
    O.scatter_shapes(G_ref, L_ref, phi_ref, g["ix_u"], g["ix_i"])
    G, L = st["Gamma_shp"].copy(), st["Lambda_shp"].copy()
    phi = np.empty((y.shape[0], 10))
    update_shapes(G, st["Gamma_rte"], L, st["Lambda_rte"], y, u, i, 0.3, 0.3, phi=phi)
    assert relerr(phi, phi_ref) < 1e-12
    assert relerr(G, G_ref) < 1e-12 and relerr(L, L_ref) < 1e-12
    # linearity in the counts (size-independent property): doubling Y doubles the sums
    G2, L2 = st["Gamma_shp"].copy(), st["Lambda_shp"].copy()
    update_shapes(G2, st["Gamma_rte"], L2, st["Lambda_rte"], 2 * y, u, i, 0.3, 0.3)
    assert relerr(G2 - 0.3, 2 * (G - 0.3)) < 1e-12


@pytest.mark.parametrize("name,upb,ipb", [("users", 20, 0), ("items", 0, 30), ("both", 20, 30)])
def test_svi_fit_vs_golden_reference(golden_svi, name, upb, ipb):
    """SVI epochs inside fit_hpf (pxi:262-377) through the module-level mirror, fp64, 8 epochs."""
    from hpfrec_b200.loops import cuda_loops_double as lp
    g = golden_svi
    Theta, Beta = np.empty((100, 10)), np.empty((100, 10))
    emp_r, emp_i = np.empty(0), np.empty(0, dtype=np.uint64)
    niter, temp, _ = lp.fit_hpf(0.3, 0.3, 1.0, 0.3, 0.3, 1.0, g["Y"], g["ix_u"].astype(np.uint64),
                                g["ix_i"].astype(np.uint64), Theta, Beta, 8, "maxiter", 0, 1e-3, upb, ipb,
                                lambda x: 1 / np.sqrt(x + 2), 0, g["st_ix_u"].astype(np.uint64), "", 123, 0, 1, 1, 0,
                                emp_r, emp_i, emp_i, 0, 1, 0)
    assert niter == 7
    got = dict(zip(STATE_KEYS, (Theta, Beta) + tuple(temp)))
    for key in STATE_KEYS:
        assert relerr(got[key], g["%s_%s" % (name, key)]) < 1e-10, key


def test_partial_fit_vs_golden_reference(golden_full, golden_pf):
    """Five mixed Cython-level partial_fit calls (pxi:423-473), in place on host arrays."""
    from hpfrec_b200.loops import cuda_loops_double as lp
    g, p = golden_full, golden_pf
    u, i, y = g["ix_u"], g["ix_i"], g["Y"]
    Theta, Beta = np.empty((100, 10)), np.empty((100, 10))
    Gs, Gr, Ls, Lr, kr, tr = lp.initialize_parameters(Theta, Beta, 123, 0.3, 0.3, 1.0, 0.3, 0.3, 1.0)
    for call in range(5):
        kind = str(p["call%d_kind" % call])
        ids = p["call%d_ids" % call]
        sel = np.isin(u, ids) if kind == "users" else np.isin(i, ids)
        ub, ib, yb = u[sel], i[sel], y[sel]
        users, items = np.unique(ub), np.unique(ib)
        lp.partial_fit(yb, ub.astype(np.uint64), ib.astype(np.uint64), Theta, Beta, Gs, Gr, Ls, Lr, kr, tr,
                       0.3, 0.3, 0.3, 0.3, 3.3, 3.3, 10, users.astype(np.uint64), items.astype(np.uint64), 0,
                       float(p["call%d_rho" % call]), 100.0 / users.shape[0], 1, kind == "users")
        for key, arr in zip(STATE_KEYS, (Theta, Beta, Gs, Gr, Ls, Lr, kr, tr)):
            assert relerr(arr, p["call%d_%s" % (call, key)]) < 1e-11, (call, key)


def test_llk_and_predict_vs_golden_reference(golden_full):
    from hpfrec_b200.loops import cuda_loops_double as lp
    g = golden_full
    u, i, y = g["ix_u"].astype(np.uint64), g["ix_i"].astype(np.uint64), g["Y"]
    T, B = g["it100_Theta"], g["it100_Beta"]
    assert abs(float(lp.calc_llk(y, u, i, T, B, 10, 1, 1)) - float(g["llk_full"])) < 1e-7
    assert abs(float(lp.calc_llk(y, u, i, T, B, 10, 1, 0)) - float(g["llk_part"])) < 1e-7
    assert relerr(lp.predict_arr(T, B, u, i, 1), g["pred"]) < 1e-13


def test_edge_cases():
    from hpfrec_b200 import _lib
    from hpfrec_b200.engine import Engine
    st = O.initialize_parameters(5, 4, 3, 1, 0.3, 1.0, 0.3, 1.0)
    # empty triple list: a full iteration degenerates to the priors
    eng = _engine_from(st, 3, np.float64)
    eng.load_coo(np.empty(0, np.int64), np.empty(0, np.int64), np.empty(0, np.float64))
    eng.step_full(1)
    out = eng.export_all()
    assert np.allclose(out["Gamma_shp"], 0.3) and np.allclose(out["Lambda_shp"], 0.3)
    ref = {k_: v.copy() for k_, v in st.items()}
    O.cavi_full_iteration(ref, np.empty(0), np.empty(0, np.int64), np.empty(0, np.int64), **HYP)
    for key in STATE_KEYS:
        assert relerr(out[key], ref[key]) < 1e-13, key
    # out-of-range index is rejected, not UB
    with pytest.raises(_lib.HPFError):
        eng.load_coo(np.array([5], np.int64), np.array([0], np.int64), np.array([1.0]))
    with pytest.raises(_lib.HPFError):
        eng.load_coo(np.array([0], np.int64), np.array([-1], np.int64), np.array([1.0]))
    eng.close()
    # step before load
    e2 = Engine(5, 4, 3, 8)
    with pytest.raises(_lib.HPFError):
        e2.step_full(1)
    e2.close()
    # one heavy row spanning many chunks + int32 indices
    nU, nI, k = 3, 2000, 6
    u = np.zeros(2000, np.int32)
    i = np.arange(2000, dtype=np.int32)
    y = np.ones(2000)
    st = O.initialize_parameters(nU, nI, k, 2, 0.3, 1.0, 0.3, 1.0)
    eng = _engine_from(st, k, np.float64, chunk=32)
    eng.load_coo(u, i, y)
    eng.step_full(3)
    out = eng.export_all()
    eng.close()
    ref = O.fit_full(y, u.astype(np.int64), i.astype(np.int64), nU, nI, k, 3, seed=2)
    for key in STATE_KEYS:
        assert relerr(out[key], ref[key]) < 1e-11, key


def test_step_granularity_and_graph_are_equivalent(golden_full):
    """step_full(5)+step_full(5) == step_full(10) == graph replay; lean iterations lose nothing."""
    import torch
    g = golden_full
    a = _fit_gpu(g["Y"], g["ix_u"], g["ix_i"], 100, 100, 10, 10, 123)
    st = O.initialize_parameters(100, 100, 10, 123, 0.3, 1.0, 0.3, 1.0)
    eng = _engine_from(st, 10, np.float64)
    eng.load_coo(g["ix_u"], g["ix_i"], g["Y"])
    eng.step_full(5)
    mid = eng.export_all()
    eng.step_full(5)
    b = eng.export_all()
    eng.close()
    assert relerr(mid["Theta"], g["it10_Theta"]) > 1e-6        # really a different iterate
    for key in STATE_KEYS:
        assert relerr(b[key], a[key]) < 1e-12, key
    stream = torch.cuda.Stream()
    eng = _engine_from(st, 10, np.float64, use_graph=1)
    eng.set_stream(stream)
    eng.load_coo(g["ix_u"], g["ix_i"], g["Y"])
    eng.step_full(10)
    c = eng.export_all()
    eng.close()
    for key in STATE_KEYS:
        assert relerr(c[key], a[key]) < 1e-12, key


def test_sharded_phases_equal_single_engine(golden_full):
    """Two user shards on one GPU driven through hpf_sweep/update_users/partials/update_items with a
    host-side sum standing in for the all-reduce == one engine (1-GPU vs N-GPU <= 1e-10)."""
    import torch
    from hpfrec_b200.dist import plan_user_shards, wrap_device_buffer
    g = golden_full
    u, i, y = g["ix_u"], g["ix_i"], g["Y"]
    its = 6
    single = _fit_gpu(y, u, i, 100, 100, 10, its, 123)
    st = O.initialize_parameters(100, 100, 10, 123, 0.3, 1.0, 0.3, 1.0)
    cuts = plan_user_shards(u, 100, 2)
    engs = []
    for r in range(2):
        lo, hi = cuts[r], cuts[r + 1]
        sel = (u >= lo) & (u < hi)
        part = dict(st)
        for key in ("Gamma_shp", "Gamma_rte", "k_rte"):
            part[key] = st[key][lo:hi]
        e = _engine_from(part, 10, np.float64)
        e.load_coo(np.ascontiguousarray(u[sel] - lo), np.ascontiguousarray(i[sel]), np.ascontiguousarray(y[sel]))
        engs.append(e)
    for _ in range(its):
        for e in engs:
            e.sweep()
            e.update_users()
        bufs = [[wrap_device_buffer(p, n, dt) for p, n, dt in
                 ((e.partials()[0], e.partials()[1], torch.float64), (e.partials()[2], e.partials()[3], torch.float64))]
                for e in engs]
        for j in range(2):
            tot = bufs[0][j] + bufs[1][j]
            bufs[0][j].copy_(tot)
            bufs[1][j].copy_(tot)
        torch.cuda.synchronize()
        for e in engs:
            e.update_items()
    outs = [e.export_all() for e in engs]
    for e in engs:
        e.close()
    for key in ("Lambda_shp", "Lambda_rte", "t_rte", "Beta"):
        assert np.array_equal(outs[0][key], outs[1][key]), key        # replicas stay bit-identical
        assert relerr(outs[0][key], single[key]) < 1e-10, key
    for key in ("Gamma_shp", "Gamma_rte", "k_rte", "Theta"):
        assert relerr(np.concatenate([outs[0][key], outs[1][key]]), single[key]) < 1e-10, key


@pytest.mark.skipif(not R.available(), reason="oracle/_ref did not travel")
def test_medium_size_vs_compiled_reference():
    """30k x 12k x 600k nnz, k=50, fp64, 5 iterations against the compiled reference run on the box."""
    mod = R.load(False)
    nU, nI, nnz, k = 30000, 12000, 600000, 50
    u, i, y = O.synth_coo(nU, nI, nnz, seed=42)
    r = R.ref_fit_hpf(mod, y, u, i, nU, nI, k, 5, seed=123, ncores=8)
    out = _fit_gpu(y, u, i, nU, nI, k, 5, 123, panel_mb=1.0)
    for key in STATE_KEYS:
        assert relerr(out[key], r[key]) < 1e-9, key


@pytest.mark.parametrize("k,dtype", [(200, np.float32), (400, np.float32), (200, np.float64)])
def test_wide_rows_vs_oracle(k, dtype):
    """Row lengths beyond 512 bytes (32-lane groups with 2 and 4 packs per lane)."""
    nU, nI, nnz = 150, 90, 3000
    u, i, y = O.synth_coo(nU, nI, nnz, seed=k)
    st0 = O.initialize_parameters(nU, nI, k, 3, 0.3, 1.0, 0.3, 1.0, dtype)
    ref = {k_: v.astype(np.float64) for k_, v in st0.items()}
    for _ in range(2):
        O.cavi_full_iteration(ref, y, u, i, **HYP)
    out = _fit_gpu(y, u, i, nU, nI, k, 2, 3, dtype=dtype, chunk=32)
    tol = 1e-11 if dtype == np.float64 else 2e-5
    for key in STATE_KEYS:
        assert relerr(out[key], ref[key]) < tol, key


# every compiled shape of sweep_rows_kernel per row class: (lanes per row, rows in flight, CTA size, CTAs per SM)
SHAPES = {(10, 8): [(0, 0, 0, 0)],
          (30, 4): [(4, 4, 256, 3), (8, 4, 256, 4), (8, 2, 256, 4), (8, 4, 128, 6), (4, 4, 256, 2), (4, 4, 128, 6)],
          (50, 4): [(8, 4, 256, 3), (8, 2, 256, 4), (8, 4, 128, 6), (8, 4, 256, 2), (8, 2, 256, 3), (16, 4, 256, 4),
                    (4, 4, 128, 4), (8, 8, 128, 4)],
          (128, 4): [(8, 4, 128, 3), (16, 4, 256, 3), (16, 2, 256, 3), (8, 2, 128, 4), (8, 4, 128, 2)],
          (50, 8): [(8, 4, 128, 3), (16, 4, 128, 4), (8, 2, 128, 3), (16, 2, 128, 4)]}


@pytest.mark.parametrize("k,dtype", [(10, np.float64), (50, np.float32), (30, np.float32), (128, np.float32),
                                     (50, np.float64)])
def test_sweep_shapes_vs_oracle(golden_full, k, dtype):
    """Every compiled lane-group shape of the sweep kernel in both production forms (shared-memory ring with
    whole-stride copies; register gathers), plus the measurement variants of the headline row class (L2-policy
    modes, prefetch, the ring without whole-stride copies): the same iterations as the oracle."""
    g = golden_full
    st0 = O.initialize_parameters(100, 100, k, 123, 0.3, 1.0, 0.3, 1.0, dtype)
    ref = {k_: v.astype(np.float64) for k_, v in st0.items()}
    for _ in range(3):
        O.cavi_full_iteration(ref, g["Y"], g["ix_u"], g["ix_i"], **HYP)
    tol = 1e-11 if dtype == np.float64 else 3e-5
    headline = (k, dtype) == (50, np.float32)
    for lpg, depth, block, minb in SHAPES[(k, np.dtype(dtype).itemsize)]:
        variants = [dict(smem_gather=1, fullrow=1, hint=0)]
        if depth <= 4:
            variants.append(dict(smem_gather=0, fullrow=0, hint=0))
            if headline:
                variants += [dict(smem_gather=0, fullrow=0, hint=1), dict(smem_gather=0, fullrow=0, hint=2),
                             dict(smem_gather=0, fullrow=1, hint=0), dict(smem_gather=0, fullrow=0, hint=0, prefetch=1)]
        if headline:
            variants += [dict(smem_gather=1, fullrow=1, hint=2), dict(smem_gather=1, fullrow=0, hint=0)]
        if k == 10:
            variants = [dict(), dict(smem_gather=0)]
        for var in variants:
            out = _fit_gpu(g["Y"], g["ix_u"], g["ix_i"], 100, 100, k, 3, 123, dtype=dtype, lpg=lpg, depth=depth, block=block,
                           minb=minb, strict=1, chunk=64, panel_mb=0.004, **var)
            for key in STATE_KEYS:
                assert relerr(out[key], ref[key]) < tol, (key, lpg, depth, block, minb, var)


def test_step_batch_ids_equals_explicit_batch(golden_full):
    """Device-assembled minibatch (ids only) == caller-assembled minibatch (triples + unique lists),
    for a user batch and an item batch, with and without blending all rates; also vs the oracle."""
    g = golden_full
    u, i, y = g["ix_u"], g["ix_i"], g["Y"]
    st = O.initialize_parameters(100, 100, 10, 123, 0.3, 1.0, 0.3, 1.0)
    for user_batch, ids in ((True, np.array([3, 50, 7, 99, 20, 21], np.int64)), (False, np.array([0, 98, 44, 45], np.int64))):
        for blend_all in (False, True):
            sel = np.isin(u, ids) if user_batch else np.isin(i, ids)
            ub, ib, yb = u[sel], i[sel], y[sel]
            users = ids if user_batch else np.unique(ub)
            items = np.unique(ib) if user_batch else ids
            e1 = _engine_from(st, 10, np.float64, panel_mb=1e9)
            e1.load_coo(u, i, y)
            e1.step_batch_ids(ids, user_batch, 0.6, 100.0 / ids.shape[0], blend_all)
            a = e1.export_all()
            e1.close()
            e2 = _engine_from(st, 10, np.float64)
            e2.step_batch(ub, ib, yb, np.ascontiguousarray(users), np.ascontiguousarray(items), user_batch, 0.6,
                          100.0 / ids.shape[0], blend_all)
            b = e2.export_all()
            e2.close()
            ref = {k_: v.copy() for k_, v in st.items()}
            O._batch_step(ref, yb, ub, ib, np.sort(users), np.sort(items), user_batch, 0.6, 100.0 / ids.shape[0],
                          0.3, 0.3, 0.3 + 10 * 0.3, 0.3 + 10 * 0.3, 0.3, 0.3, blend_all)
            for key in STATE_KEYS:
                assert relerr(a[key], b[key]) < 1e-12, (key, user_batch, blend_all)
                assert relerr(a[key], ref[key]) < 1e-11, (key, user_batch, blend_all)
    # ids-only batches need single-panel orderings
    from hpfrec_b200 import _lib
    e3 = _engine_from(st, 10, np.float64, panel_mb=0.001)
    e3.load_coo(u, i, y)
    with pytest.raises(_lib.HPFError):
        e3.step_batch_ids(np.array([1, 2], np.int64), True, 0.5, 50.0, False)
    e3.close()


def test_fp32_svi_tracks_fp64(golden_svi):
    """fp32 SVI (device-assembled batches) stays close to the fp64 reference run over 8 epochs."""
    from hpfrec_b200.loops import cuda_loops_float as lp
    g = golden_svi
    Theta, Beta = np.empty((100, 10), np.float32), np.empty((100, 10), np.float32)
    emp_r, emp_i = np.empty(0, np.float32), np.empty(0, dtype=np.uint64)
    lp.fit_hpf(0.3, 0.3, 1.0, 0.3, 0.3, 1.0, g["Y"].astype(np.float32), g["ix_u"].astype(np.uint64),
               g["ix_i"].astype(np.uint64), Theta, Beta, 8, "maxiter", 0, 1e-3, 20, 30, lambda x: 1 / np.sqrt(x + 2), 0,
               g["st_ix_u"].astype(np.uint64), "", 123, 0, 1, 1, 0, emp_r, emp_i, emp_i, 0, 1, 0)
    assert np.isfinite(Theta).all() and np.isfinite(Beta).all()
    # different random start (float32 draws) => compare the fitted model, not the iterates: predicted
    # counts of the training pairs correlate strongly with the fp64 fit's
    pred32 = np.einsum("nk,nk->n", Theta[g["ix_u"]], Beta[g["ix_i"]])
    pred64 = np.einsum("nk,nk->n", g["both_Theta"][g["ix_u"]], g["both_Beta"][g["ix_i"]])
    assert np.corrcoef(pred32, pred64)[0, 1] > 0.8


def _expected_ld(k, itemsize, align):
    """Row stride rule of hpf_create: k rounded up to 16-byte packs, then to `align` bytes but never
    beyond the next power of two of the row, then to whole 32-byte sectors."""
    pack = 16 // itemsize
    kw = -(-k // pack) * pack
    ld = kw
    if kw * itemsize > 32:
        per, cap = align // itemsize, pack
        while cap < kw:
            cap *= 2
        ld = min(-(-kw // per) * per, cap)
    per32 = 32 // itemsize
    return -(-ld // per32) * per32


@pytest.mark.parametrize("align", [32, 64, 128, 256])
@pytest.mark.parametrize("k,dtype,ld32", [(50, np.float32, 56), (10, np.float64, 12), (7, np.float64, 8),
                                          (30, np.float32, 32), (70, np.float32, 72), (3, np.float32, 8)])
def test_row_alignment_is_result_neutral(monkeypatch, golden_full, align, k, dtype, ld32):
    """HPF_ROW_ALIGN only changes the row stride of the device matrices (32 B = whole sectors, 128 B =
    whole cache lines): the stride follows the documented rule and every sweep implementation returns
    the same iteration as the oracle."""
    from hpfrec_b200.engine import Engine
    monkeypatch.setenv("HPF_ROW_ALIGN", str(align))
    probe = Engine(5, 5, k, np.dtype(dtype).itemsize)
    ld = probe.ld
    probe.close()
    assert ld == _expected_ld(k, np.dtype(dtype).itemsize, align)
    if align == 32:
        assert ld == ld32   # whole sectors: the layout of DESIGN.md section 4
    g = golden_full
    st0 = O.initialize_parameters(100, 100, k, 123, 0.3, 1.0, 0.3, 1.0, dtype)
    ref = {k_: v.astype(np.float64) for k_, v in st0.items()}
    for _ in range(3):
        O.cavi_full_iteration(ref, g["Y"], g["ix_u"], g["ix_i"], **HYP)
    tol = 1e-11 if dtype == np.float64 else 3e-5
    for sweep in (0, 1):
        out = _fit_gpu(g["Y"], g["ix_u"], g["ix_i"], 100, 100, k, 3, 123, dtype=dtype, sweep=sweep, chunk=32,
                       panel_mb=0.004)
        for key in STATE_KEYS:
            assert relerr(out[key], ref[key]) < tol, (key, sweep, align)


def test_row_alignment_minibatch_and_metrics(monkeypatch, golden_full, golden_pf):
    """The minibatch step, llk and predict kernels with cache-line-aligned rows (stride != active width)."""
    monkeypatch.setenv("HPF_ROW_ALIGN", "128")
    test_partial_fit_vs_golden_reference(golden_full, golden_pf)
    test_llk_and_predict_vs_golden_reference(golden_full)
    test_step_batch_ids_equals_explicit_batch(golden_full)


def test_engine_options_from_environment(monkeypatch, golden_full):
    """HPF_OPTIONS applies hpf_set_option defaults to every new engine; unknown names fail loudly."""
    from hpfrec_b200 import _lib
    from hpfrec_b200.engine import Engine
    g = golden_full
    monkeypatch.setenv("HPF_OPTIONS", "sweep=1,chunk=32,panel_mb=0.01")
    out = _fit_gpu(g["Y"], g["ix_u"], g["ix_i"], 100, 100, 10, 10, 123)
    for key in STATE_KEYS:
        assert relerr(out[key], g["it10_%s" % key]) < 1e-11, key
    monkeypatch.setenv("HPF_OPTIONS", "no_such_option=1")
    with pytest.raises(_lib.HPFError):
        Engine(5, 5, 4, 8)
    monkeypatch.setenv("HPF_OPTIONS", "sweep")
    with pytest.raises(_lib.HPFError):
        Engine(5, 5, 4, 8)


@pytest.mark.parametrize("k,dtype", [(4, np.float32), (7, np.float64), (10, np.float64), (30, np.float32),
                                     (50, np.float32), (50, np.float64), (64, np.float32), (100, np.float32),
                                     (128, np.float32), (128, np.float64), (200, np.float32), (400, np.float32)])
@pytest.mark.parametrize("chunk", [32, 96, 256])
def test_sweep_kernel_vs_oracle(k, dtype, chunk):
    """sweep_rows_kernel (cp.async rings three steps ahead, triples staged through shared memory, padded
    chunks) on ragged data with empty rows, chunk lengths that leave most of the last chunk as padding,
    several L2 panels: 2 iterations vs the fp64 oracle from the same start."""
    nU, nI, nnz = 500, 260, 9000
    u, i, y = O.synth_coo(nU - 40, nI - 30, nnz, seed=k + chunk)     # the last 40 users / 30 items have no data
    st0 = O.initialize_parameters(nU, nI, k, 11, 0.3, 1.0, 0.3, 1.0, dtype)
    ref = {k_: v.astype(np.float64) for k_, v in st0.items()}
    for _ in range(2):
        O.cavi_full_iteration(ref, y, u, i, **HYP)
    tol = 1e-11 if dtype == np.float64 else 2e-5
    out = _fit_gpu(y, u, i, nU, nI, k, 2, 11, dtype=dtype, sweep=0, panel_mb=0.03, chunk=chunk)
    for key in STATE_KEYS:
        assert relerr(out[key], ref[key]) < tol, key


@pytest.mark.parametrize("lpg", [0, 8])
def test_sweep_kernel_golden_trajectory(golden_full, lpg):
    """100 full-batch iterations of the README toy vs the compiled reference (generic shape and 8 lanes per row)."""
    g = golden_full
    for its, tol in ((1, 1e-12), (10, 1e-11), (100, 1e-9)):
        out = _fit_gpu(g["Y"], g["ix_u"], g["ix_i"], 100, 100, 10, its, 123, lpg=lpg, chunk=32)
        for key in STATE_KEYS:
            assert relerr(out[key], g["it%d_%s" % (its, key)]) < tol, key


@pytest.mark.parametrize("dtype,k,tol", [(np.float64, 10, 1e-12), (np.float32, 50, 3e-5), (np.float64, 50, 1e-12)])
def test_user_update_under_the_item_pass_is_result_neutral(golden_full, dtype, k, tol):
    """overlap_update: the user update on a second stream under the item-major pass, new factors into a second buffer
    (the buffers swap roles every iteration).  Same iterations as the oracle, odd and even iteration counts, mixed with
    lean / materialising calls and a minibatch step in between (which reads the CURRENT factor buffer)."""
    g = golden_full
    st0 = O.initialize_parameters(100, 100, k, 123, 0.3, 1.0, 0.3, 1.0, dtype)
    ref = {k_: v.astype(np.float64) for k_, v in st0.items()}
    eng = _engine_from(st0, k, dtype, overlap_update=148, chunk=64, panel_mb=0.01)
    eng.set_hyper(HYP["a"], HYP["a_prime"], HYP["b_prime"], HYP["c"], HYP["c_prime"], HYP["d_prime"])
    eng.load_coo(np.ascontiguousarray(g["ix_u"], np.int64), np.ascontiguousarray(g["ix_i"], np.int64),
                 np.ascontiguousarray(g["Y"], dtype))
    done = 0
    for n in (1, 2, 3, 1):
        eng.step_full(n)
        for _ in range(n):
            O.cavi_full_iteration(ref, g["Y"], g["ix_u"], g["ix_i"], **HYP)
        done += n
        out = eng.export_all()
        for key in STATE_KEYS:
            assert relerr(out[key], ref[key]) < tol, (key, done)
    # the non-overlapped path continues from the same state
    eng.set_option("overlap_update", 0)
    eng.step_full(2)
    for _ in range(2):
        O.cavi_full_iteration(ref, g["Y"], g["ix_u"], g["ix_i"], **HYP)
    out = eng.export_all()
    for key in STATE_KEYS:
        assert relerr(out[key], ref[key]) < tol, key
    eng.close()
