"""GPU tests (-m gpu) at BASELINE.json's full sizes (1M x 380K users x items, 48M nnz) through
size-independent properties of the sweep, since the CPU oracle cannot run there in seconds:

  * conservation: every phi row sums to its count, so after any iteration
        sum_j (Gamma_shp[u,j] - a)  == sum of the counts of user u      (per row, and in total)
        sum_j (Lambda_shp[i,j] - c) == sum of the counts of item i
    -- a checksum of the scatter indices and of the softmax normalisation;
  * closed forms: k_rte == a'/b' + rowsum(Theta), t_rte == c'/d' + rowsum(Beta), and the rank-one
    structure Gamma_rte[u,j] - Gamma_rte[u,0] independent of u (pxi:236);
  * two independent implementations (two-pass segmented sweep vs single-pass COO atomics) agree.

Configs: C2 (k=30), H (k=50), C3 (k=128), fp32; H also in fp64.
"""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

NU, NI, NNZ = 1_000_000, 380_000, 48_000_000
A = C = AP = CP = 0.3
BP = DP = 1.0


@pytest.fixture(scope="module")
def data():
    import torch
    import bench
    dev = torch.device("cuda", 0)
    u, i, y = bench.synth_coo_torch(NU, NI, NNZ, dev)
    u32, i32 = u.to(torch.int32).contiguous(), i.to(torch.int32).contiguous()
    ucnt = torch.zeros(NU, dtype=torch.float64, device=dev).index_add_(0, u, y.double()).cpu().numpy()
    icnt = torch.zeros(NI, dtype=torch.float64, device=dev).index_add_(0, i, y.double()).cpu().numpy()
    return u32, i32, y, ucnt, icnt


def _run(data, k, dtype, its, **opts):
    import torch
    from hpfrec_b200.engine import Engine
    from hpfrec_b200.loops import CudaLoops
    u32, i32, y, _, _ = data
    rb = np.dtype(dtype).itemsize
    loops = CudaLoops(rb == 4, device=0)
    st = loops.initialize_parameters(np.empty((NU, k), dtype), np.empty((NI, k), dtype), 123, A, AP, BP, C, CP, DP)
    eng = Engine(NU, NI, k, rb, 0)
    for name, val in opts.items():
        eng.set_option(name, val)
    eng.load_state(*st)
    eng.load_coo(u32, i32, y.to(torch.float32 if rb == 4 else torch.float64).contiguous())
    eng.step_full(its)
    out = eng.export_all()
    eng.close()
    return out


def _check_invariants(out, data, k, tol):
    _, _, _, ucnt, icnt = data
    gu = (out["Gamma_shp"].astype(np.float64) - A).sum(axis=1)
    li = (out["Lambda_shp"].astype(np.float64) - C).sum(axis=1)
    total = ucnt.sum()
    assert abs(gu.sum() - total) / total < tol / 10
    assert abs(li.sum() - total) / total < tol / 10
    assert np.max(np.abs(gu - ucnt) / np.maximum(ucnt, 1.0)) < tol
    assert np.max(np.abs(li - icnt) / np.maximum(icnt, 1.0)) < tol
    theta = out["Theta"].astype(np.float64)
    beta = out["Beta"].astype(np.float64)
    assert np.isfinite(theta).all() and np.isfinite(beta).all()
    assert np.max(np.abs(out["k_rte"][:, 0] - (AP / BP + theta.sum(axis=1))) / out["k_rte"][:, 0]) < tol
    assert np.max(np.abs(out["t_rte"][:, 0] - (CP / DP + beta.sum(axis=1))) / out["t_rte"][:, 0]) < tol
    gr = out["Gamma_rte"].astype(np.float64)
    spread = (gr[:, 1:] - gr[:, :1]).std(axis=0) / np.abs(gr).mean()
    assert spread.max() < tol
    assert np.allclose(theta, out["Gamma_shp"].astype(np.float64) / gr, rtol=tol)


@pytest.mark.parametrize("k", [30, 50, 128])
def test_fp32_full_size_invariants(data, k):
    out = _run(data, k, np.float32, 3)
    _check_invariants(out, data, k, 2e-4)


def test_fp64_full_size_invariants(data):
    out = _run(data, 50, np.float64, 2)
    _check_invariants(out, data, 50, 1e-10)


def test_two_sweep_implementations_agree_at_full_size(data):
    a = _run(data, 50, np.float32, 2, sweep=0)
    b = _run(data, 50, np.float32, 2, sweep=1)
    for key in ("Theta", "Beta"):
        num = np.abs(a[key].astype(np.float64) - b[key])
        assert np.max(num / np.maximum(np.abs(b[key]), 1e-3)) < 5e-4, key
