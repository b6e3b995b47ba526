#!/usr/bin/env python3
"""Generates tests/golden/*.npz from the UNMODIFIED compiled reference (oracle/_ref, built from
/root/reference by oracle/build_ref.py).  Run in the build container:

    python oracle/build_ref.py && python tests/golden/make_golden.py

The reference ships no golden vectors of its own, so these are "outputs of the
reference itself run here".  Versions that pin them (recorded in every file): numpy, scipy, cython.
Everything is fp64, ncores=1, allow_inconsistent_math=False (par_sh=0) unless noted.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import hpf_oracle as O  # noqa: E402
from oracle import ref_loader as R  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))
KEYS = ("Theta", "Beta", "Gamma_shp", "Gamma_rte", "Lambda_shp", "Lambda_rte", "k_rte", "t_rte")


def versions():
    import scipy, Cython
    return "numpy %s scipy %s cython %s" % (np.__version__, scipy.__version__, Cython.__version__)


def toy():
    df = O.readme_toy()
    return (df.UserId.to_numpy().astype(np.int64), df.ItemId.to_numpy().astype(np.int64),
            df.Count.to_numpy().astype(np.float64))


def main():
    mod = R.load(False)
    modf = R.load(True)
    assert mod is not None, "build oracle/_ref first"
    u, i, y = toy()
    nU = nI = 100
    k = 10
    ver = versions()

    # --- full batch trajectory on the README toy (config C1) ---------------------------------
    full = {"versions": ver, "ix_u": u, "ix_i": i, "Y": y}
    for its in (1, 2, 10, 100):
        r = R.ref_fit_hpf(mod, y, u, i, nU, nI, k, its, seed=123)
        for key in KEYS:
            full["it%d_%s" % (its, key)] = r[key]
        full["it%d_niter" % its] = r["niter"]
    # float build, 1 and 10 iterations (the fp32 engine is gated on single-sweep parity)
    for its in (1, 10):
        r = R.ref_fit_hpf(modf, y.astype(np.float32), u, i, nU, nI, k, its, seed=123)
        for key in ("Theta", "Beta"):
            full["f32_it%d_%s" % (its, key)] = r[key]
    # llk / predictions at the 100-iteration state
    r = R.ref_fit_hpf(mod, y, u, i, nU, nI, k, 100, seed=123)
    ind = mod.obj_ind_type
    full["llk_full"] = np.float64(mod.calc_llk(y, u.astype(ind), i.astype(ind), r["Theta"], r["Beta"], k, 1, 1))
    full["llk_part"] = np.float64(mod.calc_llk(y, u.astype(ind), i.astype(ind), r["Theta"], r["Beta"], k, 1, 0))
    full["pred"] = mod.predict_arr(r["Theta"], r["Beta"], u.astype(ind), i.astype(ind), 1)
    np.savez_compressed(os.path.join(OUT, "toy_full.npz"), **full)

    # --- odd shapes: k not a multiple of the pack, empty users/items, ragged degrees ------------
    rng = np.random.default_rng(7)
    nU2, nI2, k2 = 37, 53, 7
    uu = rng.integers(0, nU2 - 3, size=400)      # last 3 users have no data
    ii = rng.integers(2, nI2, size=400)          # first 2 items have no data
    key = np.unique(uu * nI2 + ii)
    uu, ii = key // nI2, key % nI2
    perm = rng.permutation(uu.shape[0])
    uu, ii = uu[perm], ii[perm]
    yy = (1 + rng.poisson(1.5, size=uu.shape[0])).astype(np.float64)
    odd = {"versions": ver, "ix_u": uu, "ix_i": ii, "Y": yy, "nU": nU2, "nI": nI2, "k": k2}
    for its in (1, 25):
        r = R.ref_fit_hpf(mod, yy, uu, ii, nU2, nI2, k2, its, seed=5, a=0.5, a_prime=0.4, b_prime=1.3,
                          c=0.6, c_prime=0.2, d_prime=0.8)
        for key_ in KEYS:
            odd["it%d_%s" % (its, key_)] = r[key_]
    np.savez_compressed(os.path.join(OUT, "odd_full.npz"), **odd)

    # --- SVI inside fit_hpf (sorted by user; ncores=1) -----------------------------------------
    order = np.argsort(u, kind="stable")
    us, is_, ys = u[order], i[order], y[order]
    st = np.zeros(nU + 1, dtype=np.int64)
    np.add.at(st, us + 1, 1)
    st = np.cumsum(st)
    svi = {"versions": ver, "ix_u": us, "ix_i": is_, "Y": ys, "st_ix_u": st}
    for name, upb, ipb in (("users", 20, 0), ("items", 0, 30), ("both", 20, 30)):
        r = R.ref_fit_hpf(mod, ys, us, is_, nU, nI, k, 8, seed=123, users_per_batch=upb, items_per_batch=ipb,
                          st_ix_u=st, par_sh=1, ncores=1)
        for key_ in KEYS:
            svi["%s_%s" % (name, key_)] = r[key_]
    np.savez_compressed(os.path.join(OUT, "toy_svi.npz"), **svi)

    # --- Cython-level partial_fit: five mixed calls on a fresh state ---------------------------
    Theta = np.empty((nU, k))
    Beta = np.empty((nI, k))
    Gs, Gr, Ls, Lr, kr, tr = mod.initialize_parameters(Theta, Beta, 123, 0.3, 0.3, 1.0, 0.3, 0.3, 1.0)
    pf = {"versions": ver}
    rngb = np.random.default_rng(11)
    a = c = 0.3
    k_shp = 0.3 + k * 0.3
    t_shp = 0.3 + k * 0.3
    for call, kind in enumerate(("users", "items", "users", "users", "items")):
        if kind == "users":
            ids = np.unique(rngb.integers(0, nU, size=20))
            sel = np.isin(u, ids)
        else:
            ids = np.unique(rngb.integers(0, nI, size=25))
            sel = np.isin(i, ids)
        ub, ib, yb = u[sel], i[sel], y[sel]
        users, items = np.unique(ub), np.unique(ib)
        rho = 1.0 if call == 0 else 1 / np.sqrt(call + 2)
        mult = float(nU) / users.shape[0]
        mod.partial_fit(yb, ub.astype(ind), ib.astype(ind), Theta, Beta, Gs, Gr, Ls, Lr, kr, tr,
                        0.3 / 1.0, 0.3 / 1.0, a, c, k_shp, t_shp, k, users.astype(ind), items.astype(ind), 0,
                        rho, mult, 1, kind == "users")
        pf["call%d_kind" % call] = kind
        pf["call%d_ids" % call] = ids
        pf["call%d_rho" % call] = rho
        for key_, arr in zip(KEYS, (Theta, Beta, Gs, Gr, Ls, Lr, kr, tr)):
            pf["call%d_%s" % (call, key_)] = arr.copy()
    np.savez_compressed(os.path.join(OUT, "toy_partial_fit.npz"), **pf)
    print("golden files written to", OUT, "|", ver)


if __name__ == "__main__" and "--api" not in sys.argv and "--foldin" not in sys.argv:
    main()


def load_reference_class():
    """The reference's own `HPF` class (hpfrec/__init__.py) imported from /root/reference, with its
    compiled submodules resolved from oracle/_ref (build container only)."""
    import importlib.util
    import types
    ref_pkg = "/root/reference/hpfrec"
    so_dir = os.path.join(ROOT, "oracle", "_ref", "hpfrec")
    R.load(False)  # puts oracle/_ref on sys.path and imports hpfrec.cython_loops_double as a namespace pkg
    for name in list(sys.modules):
        if name == "hpfrec" or name.startswith("hpfrec."):
            del sys.modules[name]
    chk = types.ModuleType("hpfrec._check_openmp")
    chk.get = lambda: 1
    spec = importlib.util.spec_from_file_location("hpfrec", os.path.join(ref_pkg, "__init__.py"),
                                                  submodule_search_locations=[ref_pkg, so_dir])
    mod = importlib.util.module_from_spec(spec)
    sys.modules["hpfrec"] = mod
    sys.modules["hpfrec._check_openmp"] = chk
    spec.loader.exec_module(mod)
    return mod.HPF


def api_golden():
    """README flow through the reference's Python class (fp64, fixed seeds): fit with reindexing,
    predict / topN / eval_llk, train-llk stopping, partial_fit sequence, predict_factors."""
    import warnings
    import pandas as pd
    HPF = load_reference_class()
    df = O.readme_toy()
    out = {"versions": versions()}
    warnings.simplefilter("ignore")
    # string ids so that reindexing really happens
    dfs = df.copy()
    dfs["UserId"] = "u" + dfs["UserId"].astype(str)
    dfs["ItemId"] = "i" + dfs["ItemId"].astype(str)
    m = HPF(k=8, use_float=False, random_seed=7, maxiter=30, verbose=False, ncores=1, check_every=None)
    m.fit(dfs.copy())
    out["fit_Theta"], out["fit_Beta"] = m.Theta, m.Beta
    out["fit_user_mapping"] = np.array(m.user_mapping_, dtype="U8")
    out["fit_item_mapping"] = np.array(m.item_mapping_, dtype="U8")
    out["fit_niter"] = m.niter
    out["pred_single"] = m.predict(user="u10", item="i11")
    out["pred_many"] = m.predict(user=["u10", "u11", "u12", "zz"], item=["i4", "i5", "i6", "i1"])
    out["topn_seen"] = np.array(m.topN(user="u10", n=10, exclude_seen=True), dtype="U8")
    out["topn_all"] = np.array(m.topN(user="u10", n=10, exclude_seen=False), dtype="U8")
    out["topn_pool"] = np.array(m.topN(user="u10", n=3, exclude_seen=False,
                                       items_pool=np.array(["i1", "i2", "i3", "i4", "i50"])), dtype="U8")
    llk = m.eval_llk(dfs.copy(), full_llk=True)
    out["eval_llk"], out["eval_nobs"] = np.float64(llk["llk"]), llk["nobs"]
    # train-llk stopping criterion (niter tells where it stopped)
    m2 = HPF(k=8, use_float=False, random_seed=7, maxiter=100, verbose=False, ncores=1, stop_crit="train-llk",
             check_every=5, stop_thr=1e-3, reindex=False)
    m2.fit(df.copy())
    out["llkstop_niter"], out["llkstop_Theta"] = m2.niter, m2.Theta
    # diff-norm criterion
    m3 = HPF(k=8, use_float=False, random_seed=7, maxiter=100, verbose=False, ncores=1, stop_crit="diff-norm",
             check_every=5, stop_thr=0.5, reindex=False)
    m3.fit(df.copy())
    out["normstop_niter"], out["normstop_Theta"] = m3.niter, m3.Theta
    # public partial_fit sequence on a fresh object (README.md:116-123)
    np.random.seed(3)
    m4 = HPF(k=8, use_float=False, random_seed=7, reindex=False, keep_data=False, verbose=False, ncores=1)
    batches = [np.unique(np.random.randint(100, size=20)) for _ in range(3)]
    m4.partial_fit(df.loc[df.UserId.isin(batches[0])].copy(), nusers=100, nitems=100)
    m4.partial_fit(df.loc[df.UserId.isin(batches[1])].copy())
    m4.partial_fit(df.loc[df.ItemId.isin(batches[2])].copy(), batch_type="items")
    for j, b in enumerate(batches):
        out["pf_batch%d" % j] = b
    out["pf_Theta"], out["pf_Beta"], out["pf_k_rte"], out["pf_niter"] = m4.Theta, m4.Beta, m4.k_rte, m4.niter
    # predict_factors for a new user (README.md:135-143), on the fitted reindex=False model m2
    np.random.seed(2)
    new = pd.DataFrame({"ItemId": np.random.choice(np.arange(100), size=20, replace=False),
                        "Count": np.random.gamma(1, 1, size=20).astype("int32")})
    new = new.loc[new.Count > 0].reset_index(drop=True)
    out["new_items"], out["new_counts"] = new.ItemId.to_numpy(), new.Count.to_numpy()
    out["new_factors"] = m2.predict_factors(new.copy(), maxiter=10, random_seed=1)
    # add_user (hpfrec/__init__.py:1060-1196) on the same fitted model: a NEW user from the same data.  The reference
    # passes cast_int(stop_thr) = 0 to calc_user_factors on this path (init:1153), i.e. it always runs `maxiter` iterations.
    m2.add_user(user_id=100, counts_df=new.copy(), update_existing=False, maxiter=10, random_seed=1)
    out["adduser_Theta_last"] = m2.Theta[-1].copy()
    out["adduser_nusers"] = m2.nusers
    out["adduser_n_seen_last"] = int(m2._n_seen_by_user[-1])
    out["adduser_st_ix_last"] = int(m2._st_ix_user[-1])
    out["adduser_seen_tail"] = np.array(m2.seen[int(m2._st_ix_user[-1]):], dtype=np.int64)
    out["adduser_topn"] = np.array(m2.topN(user=100, n=5, exclude_seen=True), dtype=np.int64)
    # SVI through the class (ncores=1 for determinism)
    m5 = HPF(k=8, use_float=False, random_seed=7, maxiter=6, verbose=False, ncores=1, users_per_batch=20,
             items_per_batch=30, reindex=False, check_every=None)
    m5.fit(df.copy())
    out["svi_Theta"], out["svi_Beta"] = m5.Theta, m5.Beta
    np.savez_compressed(os.path.join(OUT, "toy_api.npz"), **out)
    print("toy_api.npz written")


def foldin_golden():
    """Fold-in of MANY new users through the reference's own predict_factors (hpfrec/__init__.py:989-1058 ->
    calc_user_factors, pxi:476-520), one call per user as the reference does it: the golden for
    HPF.predict_factors_batch.  Two stopping thresholds: the default, and a loose one that stops users at different
    iterations."""
    import warnings
    import pandas as pd
    HPF = load_reference_class()
    df = O.readme_toy()
    warnings.simplefilter("ignore")
    m = HPF(k=8, use_float=False, random_seed=7, maxiter=100, verbose=False, ncores=1, stop_crit="train-llk",
            check_every=5, stop_thr=1e-3, reindex=False)
    m.fit(df.copy())
    rng = np.random.default_rng(5)
    users, items, counts = [], [], []
    for u in range(30):
        n = int(rng.integers(1, 30))
        it = rng.choice(100, size=n, replace=False)
        users.append(np.full(n, u)); items.append(it); counts.append(rng.integers(1, 6, size=n).astype(float))
    users, items, counts = np.concatenate(users), np.concatenate(items), np.concatenate(counts)
    out = {"versions": versions(), "fit_Theta": m.Theta, "users": users, "items": items, "counts": counts}
    for tag, thr in (("default", 1e-3), ("loose", 3e-2)):
        th = np.empty((30, 8))
        for u in range(30):
            sel = users == u
            th[u] = m.predict_factors(pd.DataFrame({"ItemId": items[sel], "Count": counts[sel]}), maxiter=10,
                                      random_seed=1, stop_thr=thr)
        out["theta_" + tag] = th
    np.savez_compressed(os.path.join(OUT, "toy_foldin.npz"), **out)
    print("toy_foldin.npz written")


if __name__ == "__main__" and "--api" in sys.argv:
    api_golden()
if __name__ == "__main__" and "--foldin" in sys.argv:
    foldin_golden()
