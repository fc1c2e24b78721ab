"""GPU parity of pmmh_flps_sv_corr against the CPU oracle and the reference's golden vectors.

Tolerances (conditional on identical ancestors, which is asserted bit-exactly):
  log-likelihood   rel <= 1e-10        (north_star / SURVEY 8d)
  gradient         <= 1e-9 * max|g|
  Hessian pieces   <= 1e-8 * max|H|
  particles        rel <= 1e-12
"""
import numpy as np
import pytest

import golden_inputs as gi
from helpers import first_mismatch_step, relerr, to_time_major

pytestmark = pytest.mark.gpu

LL_TOL = 1e-10
GRAD_TOL = 1e-9
HESS_TOL = 1e-8


def _run_dev(dev, obs, params, rvr, rvp, n, nobs, lag, hess, store_history=True, ctas=0):
    import torch
    from pmmh_qn_b200 import kernels as K
    u = torch.from_numpy(to_time_major(rvp, n, nobs)).to(dev)
    out = K.flps_sv_corr(torch.from_numpy(obs).to(dev), torch.from_numpy(params).to(dev),
                         torch.from_numpy(rvr[:nobs].copy()).to(dev), u, lag=lag,
                         compute_hessian=bool(hess), store_history=store_history,
                         ctas_per_problem=ctas)
    torch.cuda.synchronize()
    return {k: v.cpu().numpy() for k, v in out.items() if not k.startswith("_")}


def _compare(res, ref, hess, check_history=True):
    assert int(res["diag"][0, 2]) == 0, "kernel reported a degenerate cloud"
    if check_history:
        step = first_mismatch_step(res["A"][0][1:], ref["A"][1:])
        assert step is None, "ancestors differ first at time %d" % (step + 1)
        assert relerr(res["X"][0], ref["X"]) <= 1e-12
    assert abs(res["log_like"][0] - ref["log_like"]) <= LL_TOL * abs(ref["log_like"])
    assert relerr(res["filt"][0], ref["filt"]) <= 1e-10
    assert relerr(res["smo"][0], ref["smo"]) <= 1e-10
    assert relerr(res["traj"][0], ref["traj"]) <= 1e-12
    g, gr = res["gradient"][0], ref["gradient"]
    assert np.max(np.abs(g - gr)) <= GRAD_TOL * np.max(np.abs(gr))
    if hess:
        for k in ("hess1", "hess2"):
            h, hr = res[k][0], ref[k]
            assert np.max(np.abs(h - hr)) <= HESS_TOL * max(np.max(np.abs(hr)), 1e-300), k


@pytest.mark.parametrize("n,nobs,lag", [(75, 361, 10), (37, 50, 10), (200, 120, 10), (64, 40, 4)])
@pytest.mark.parametrize("seed", [0, 1, 2])
@pytest.mark.parametrize("hess", [0, 1])
def test_flps_vs_oracle_small(cuda_dev, n, nobs, lag, seed, hess):
    import oracle
    obs, params, rvr, rvp = gi.sv_inputs(n, nobs, seed)
    ref = oracle.flps_sv_corr(obs, params, rvr, rvp, n, lag, hess, dumps=True)
    res = _run_dev(cuda_dev, obs, params, rvr, rvp, n, nobs, lag, hess)
    _compare(res, ref, hess)


@pytest.mark.parametrize("n,ctas", [(1024, 0), (4096, 1), (4096, 4), (4096, 148), (5000, 7)])
@pytest.mark.parametrize("hess", [0, 1])
def test_flps_vs_oracle_T1000(cuda_dev, n, ctas, hess):
    """BASELINE config sizes (T=1000) incl. multi-CTA teams; oracle runs in ~1 s here."""
    import oracle
    nobs, lag = 1001, 10
    obs, params, rvr, rvp = gi.sv_inputs(n, nobs, seed=0)
    ref = oracle.flps_sv_corr(obs, params, rvr, rvp, n, lag, hess, dumps=True)
    res = _run_dev(cuda_dev, obs, params, rvr, rvp, n, nobs, lag, hess, ctas=ctas)
    _compare(res, ref, hess)


def test_flps_vs_golden(cuda_dev, golden):
    """Against outputs of the reference's own compiled flps_sv_corr (tests/golden)."""
    g = golden["sv_kernels"]
    for (n, nobs, lag, seeds) in gi.SV_KERNEL_CASES:
        for seed in seeds:
            obs, params, rvr, rvp = gi.sv_inputs(n, nobs, seed)
            for hess in (0, 1):
                pre = "flps_n%d_t%d_l%d_s%d_h%d_" % (n, nobs, lag, seed, hess)
                ref = {k: g[pre + k] for k in ("filt", "smo", "gradient", "traj", "hess1", "hess2")}
                ref["log_like"] = float(g[pre + "log_like"])
                ref["gradient"] = ref["gradient"].reshape(4, nobs)
                ref["hess1"] = ref["hess1"].reshape(4, 4)
                ref["hess2"] = ref["hess2"].reshape(4, 4)
                res = _run_dev(cuda_dev, obs, params, rvr, rvp, n, nobs, lag, hess,
                               store_history=False)
                _compare(res, ref, hess, check_history=False)


def test_flps_ring_equals_history(cuda_dev):
    """The ring-only mode must give the same outputs as the full-history mode."""
    n, nobs, lag = 600, 200, 10
    obs, params, rvr, rvp = gi.sv_inputs(n, nobs, 1)
    a = _run_dev(cuda_dev, obs, params, rvr, rvp, n, nobs, lag, 1, store_history=True)
    b = _run_dev(cuda_dev, obs, params, rvr, rvp, n, nobs, lag, 1, store_history=False)
    for k in ("log_like", "filt", "smo", "gradient", "traj", "hess1", "hess2"):
        assert np.array_equal(a[k], b[k]), k


def test_flps_batch_matches_single(cuda_dev):
    """B independent (params, u) problems in one launch == B single launches."""
    import torch
    from pmmh_qn_b200 import kernels as K
    n, nobs, lag, B = 300, 150, 10, 9
    obs = gi.sv_obs(nobs)
    rs = np.random.RandomState(5)
    params = np.array(gi.SV_PARAM_SETS[0]) + 0.02 * rs.normal(size=(B, 4))
    rvs = rs.normal(size=(B, nobs, n + 1))
    rvr = np.zeros((B, nobs))
    u = np.zeros((B, nobs, n))
    for b in range(B):
        r, p = gi.split_particle(rvs[b], nobs)
        rvr[b] = r
        u[b] = to_time_major(p, n, nobs)
    dev = cuda_dev
    out = K.flps_sv_corr(torch.from_numpy(obs).to(dev), torch.from_numpy(params).to(dev),
                         torch.from_numpy(rvr).to(dev), torch.from_numpy(u).to(dev), lag=lag,
                         compute_hessian=True)
    torch.cuda.synchronize()
    for b in range(B):
        single = K.flps_sv_corr(torch.from_numpy(obs).to(dev), torch.from_numpy(params[b:b + 1]).to(dev),
                                torch.from_numpy(rvr[b:b + 1]).to(dev),
                                torch.from_numpy(u[b:b + 1]).to(dev), lag=lag, compute_hessian=True)
        torch.cuda.synchronize()
        for k in ("log_like", "filt", "smo", "gradient", "hess1", "hess2", "traj"):
            assert torch.equal(out[k][b], single[k][0]), (b, k)


def test_flps_degenerate_cloud_reports_status(cuda_dev):
    """sigma_v = 0 collapses the cloud: the kernel must finish and say so, not hang."""
    import torch
    from pmmh_qn_b200 import kernels as K
    n, nobs = 8192, 20
    obs = gi.sv_obs(nobs)
    params = np.array([0.2, 0.9, 0.0, -0.5])
    rvr, rvp = gi.split_particle(gi.sv_rvs(n, nobs, 0), nobs)
    dev = cuda_dev
    out = K.flps_sv_corr(torch.from_numpy(obs).to(dev), torch.from_numpy(params).to(dev),
                         torch.from_numpy(rvr).to(dev),
                         torch.from_numpy(to_time_major(rvp, n, nobs)).to(dev), lag=10)
    torch.cuda.synchronize()
    assert int(out["diag"][0, 2]) == 1
    assert not np.isfinite(float(out["log_like"][0]))
