"""CPU tests (-m "not gpu") of the multi-GPU host logic: nnz-balanced user sharding and the
per-iteration exchange driver (hpfrec_b200/dist.py) run with world_size 2 over gloo.

No GPU here, so the ranks drive a *test double* of the engine (numpy, built on the oracle -- test
infrastructure only) through exactly the phase protocol the real Engine exposes
(sweep / update_users / <all-reduce of item sums and Theta column sums> / update_items), and the
result must equal the single-process oracle fit."""
import os
import socket

import numpy as np
import pytest

from conftest import STATE_KEYS, relerr
from oracle import hpf_oracle as O
from hpfrec_b200 import dist as hdist

HYP = dict(a=0.3, a_prime=0.3, b_prime=1.0, c=0.3, c_prime=0.3, d_prime=1.0)


class FakeShardEngine:
    """numpy stand-in implementing the Engine phase protocol for one user shard (fp64)."""

    def __init__(self, st, u, i, y, k):
        import torch
        self.st, self.u, self.i, self.y, self.k = st, u, i, y, k
        nI = st["Lambda_shp"].shape[0]
        self.item_sums = torch.zeros(nI * k, dtype=torch.float64)
        self.theta_colsum = torch.zeros(k, dtype=torch.float64)
        self.beta_colsum = (st["Lambda_shp"] / st["Lambda_rte"]).sum(axis=0)

    def sweep(self):
        st = self.st
        phi = O.phi_rows(st["Gamma_shp"], st["Gamma_rte"], st["Lambda_shp"], st["Lambda_rte"], self.y, self.u,
                         self.i, True)
        self._usum = np.zeros_like(st["Gamma_shp"])
        isum = np.zeros_like(st["Lambda_shp"])
        O.scatter_shapes(self._usum, isum, phi, self.u, self.i)
        self.item_sums.copy_(__import__("torch").from_numpy(isum.reshape(-1)))

    def update_users(self):
        st, k = self.st, self.k
        st["Gamma_rte"] = (HYP["a_prime"] + k * HYP["a"]) / st["k_rte"] + self.beta_colsum[None, :]
        st["Gamma_shp"] = HYP["a"] + self._usum
        theta = st["Gamma_shp"] / st["Gamma_rte"]
        st["k_rte"] = HYP["a_prime"] / HYP["b_prime"] + theta.sum(axis=1, keepdims=True)
        self.theta_colsum.copy_(__import__("torch").from_numpy(theta.sum(axis=0)))
        st["Theta"] = theta

    def update_items(self):
        st, k = self.st, self.k
        tsum = self.theta_colsum.numpy()
        st["Lambda_shp"] = HYP["c"] + self.item_sums.numpy().reshape(-1, k)      # prior added once, after the reduce
        st["Lambda_rte"] = (HYP["c_prime"] + k * HYP["c"]) / st["t_rte"] + tsum[None, :]
        beta = st["Lambda_shp"] / st["Lambda_rte"]
        st["t_rte"] = HYP["c_prime"] / HYP["d_prime"] + beta.sum(axis=1, keepdims=True)
        self.beta_colsum = beta.sum(axis=0)
        st["Beta"] = beta


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, its, ret):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        df = O.readme_toy()
        u = df.UserId.to_numpy().astype(np.int64)
        i = df.ItemId.to_numpy().astype(np.int64)
        y = df.Count.to_numpy().astype(np.float64)
        k = 10
        st = O.initialize_parameters(100, 100, k, 123, 0.3, 1.0, 0.3, 1.0)
        cuts = hdist.plan_user_shards(u, 100, world)
        lo, hi = cuts[rank], cuts[rank + 1]
        lu, li, ly = hdist.shard_triples(u, i, y, lo, hi)
        part = {key: (val[lo:hi].copy() if key in ("Gamma_shp", "Gamma_rte", "k_rte", "Theta") else val.copy())
                for key, val in st.items()}
        eng = FakeShardEngine(part, lu, li, ly, k)
        hdist.run_sharded_iterations(eng, its, partial_tensors=(eng.item_sums, eng.theta_colsum))
        ret[rank] = (lo, hi, {key: part[key] for key in STATE_KEYS})
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_sharded_driver_matches_single_process(world):
    import torch.multiprocessing as mp
    its = 5
    port = _free_port()
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, port, its, ret), nprocs=world, join=True)
    df = O.readme_toy()
    ref = O.fit_full(df.Count.to_numpy().astype(np.float64), df.UserId.to_numpy().astype(np.int64),
                     df.ItemId.to_numpy().astype(np.int64), 100, 100, 10, its, seed=123, sum_exp_trick=True)
    parts = [ret[r] for r in range(world)]
    assert parts[0][0] == 0 and parts[-1][1] == 100
    for key in ("Gamma_shp", "Gamma_rte", "k_rte", "Theta"):
        got = np.concatenate([p[2][key] for p in parts])
        assert relerr(got, ref[key]) < 1e-11, key
    for key in ("Lambda_shp", "Lambda_rte", "t_rte", "Beta"):
        for p in parts:
            assert relerr(p[2][key], ref[key]) < 1e-11, key
            assert np.array_equal(p[2][key], parts[0][2][key])       # replicas identical after all-reduce


def test_plan_user_shards_balances_nnz():
    rng = np.random.default_rng(0)
    deg = rng.zipf(1.5, size=5000).clip(max=3000)
    u = np.repeat(np.arange(5000), deg)
    for world in (1, 2, 4, 8):
        cuts = hdist.plan_user_shards(u, 5000, world)
        assert cuts[0] == 0 and cuts[-1] == 5000 and len(cuts) == world + 1
        assert all(cuts[j] <= cuts[j + 1] for j in range(world))
        sizes = [int(((u >= cuts[r]) & (u < cuts[r + 1])).sum()) for r in range(world)]
        assert sum(sizes) == u.shape[0]
        assert max(sizes) - min(sizes) <= 2 * deg.max() + 2    # each cut is within one user of the ideal
    # torch input gives the same cuts
    import torch
    assert hdist.plan_user_shards(torch.from_numpy(u), 5000, 4) == hdist.plan_user_shards(u, 5000, 4)
    # degenerate: more ranks than users with data
    cuts = hdist.plan_user_shards(np.array([0, 0, 0]), 2, 4)
    assert cuts[0] == 0 and cuts[-1] == 2 and all(cuts[j] <= cuts[j + 1] for j in range(4))


def test_shard_triples_localises_user_ids():
    u = np.array([0, 5, 9, 5, 3])
    i = np.array([1, 2, 3, 4, 5])
    y = np.array([1., 2., 3., 4., 5.])
    lu, li, ly = hdist.shard_triples(u, i, y, 4, 10)
    assert lu.tolist() == [1, 5, 1] and li.tolist() == [2, 3, 4] and ly.tolist() == [2., 3., 4.]


def test_resolve_overlap_schedule(monkeypatch):
    """Exchange schedule of the fused modes: auto by shard size, explicit values of HPF_EXCHANGE_OVERLAP."""
    from hpfrec_b200.dist import resolve_overlap
    monkeypatch.delenv("HPF_EXCHANGE_OVERLAP", raising=False)
    assert resolve_overlap(125_000) == "update" and resolve_overlap(1_250_000) == "update"
    assert resolve_overlap(15_000) is True
    for env, want in (("1", True), ("0", False), ("update", "update"), ("auto", True)):
        monkeypatch.setenv("HPF_EXCHANGE_OVERLAP", env)
        assert resolve_overlap(15_000) == want
