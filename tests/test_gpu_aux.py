"""GPU parity of the smaller kernels: u layout change, Phi, Crank-Nicolson, the random-effects
importance sampler, the subsampling sort + stratified indices and the logistic gather-reduce."""
import numpy as np
import pytest
from scipy.stats import norm

import golden_inputs as gi
from helpers import relerr, to_time_major

pytestmark = pytest.mark.gpu


def _t(a, dev):
    import torch
    return torch.from_numpy(np.ascontiguousarray(a)).to(dev)


# ---------------------------------------------------------------- layout / Phi / CN
@pytest.mark.parametrize("n,nobs,batch", [(75, 361, 1), (1000, 33, 3), (33, 1000, 2), (1, 2, 1)])
def test_split_rvs_matches_reference_flat_split(cuda_dev, n, nobs, batch):
    """particle_methods/cython.py:89-91: first NOBS FLAT entries, then rvp[i + j*NOBS]."""
    from pmmh_qn_b200 import kernels as K
    rs = np.random.RandomState(3)
    rvs = rs.normal(size=(batch, nobs, n + 1))
    r_raw, u = K.split_rvs(_t(rvs, cuda_dev), nobs, n)
    for b in range(batch):
        flat = rvs[b].flatten()
        assert np.array_equal(r_raw[b].cpu().numpy(), flat[:nobs])
        assert np.array_equal(u[b].cpu().numpy(), to_time_major(flat[nobs:], n, nobs))


def test_norm_cdf_matches_scipy(cuda_dev):
    from pmmh_qn_b200 import kernels as K
    x = np.concatenate([np.random.RandomState(0).normal(size=100000), np.linspace(-30, 9, 2001)])
    got = K.norm_cdf(_t(x, cuda_dev)).cpu().numpy()
    want = norm.cdf(x)
    # both sides round x / sqrt(2) once; in the lower tail that error is amplified by ~x^2
    tol = 2.3e-16 * (8.0 + x * x)
    assert np.all(np.abs(got - want) <= tol * want)


def test_crank_nicolson_with_supplied_noise_is_bit_exact(cuda_dev):
    """parameter/mcmc/base_class.py:231-233 with xi supplied."""
    import oracle
    from pmmh_qn_b200 import kernels as K
    rs = np.random.RandomState(1)
    for shape, sigma_u in (((361, 76), 0.5), ((100, 101), 0.05), ((5500,), 0.05), ((7,), 0.999)):
        u = rs.normal(size=shape)
        xi = rs.normal(size=shape)
        got = K.crank_nicolson(_t(u, cuda_dev), sigma_u, xi=_t(xi, cuda_dev)).cpu().numpy()
        assert np.array_equal(got, oracle.crank_nicolson(u, xi, sigma_u))


def test_crank_nicolson_philox_statistics_and_reproducibility(cuda_dev):
    import torch
    from pmmh_qn_b200 import kernels as K
    n = 1 << 22
    u = torch.zeros(n + 1, dtype=torch.float64, device=cuda_dev)   # odd length on purpose
    a = K.crank_nicolson(u, 1.0, seed=42, philox_offset=7)
    b = K.crank_nicolson(u, 1.0, seed=42, philox_offset=7)
    c = K.crank_nicolson(u, 1.0, seed=43, philox_offset=7)
    assert torch.equal(a, b) and not torch.equal(a, c)
    x = a.cpu().numpy()
    assert abs(x.mean()) < 5.0 / np.sqrt(n)
    assert abs(x.var() - 1.0) < 5.0 * np.sqrt(2.0 / n)
    assert abs(np.mean(x ** 4) - 3.0) < 0.05
    assert abs(np.corrcoef(x[:-1:2], x[1::2])[0, 1]) < 5.0 / np.sqrt(n / 2)
    # stationarity of the CN move: u ~ N(0,1) stays N(0,1)
    u0 = torch.from_numpy(np.random.RandomState(0).normal(size=n)).to(cuda_dev)
    u1 = K.crank_nicolson(u0, 0.5, seed=1).cpu().numpy()
    assert abs(u1.var() - 1.0) < 5.0 * np.sqrt(2.0 / n)
    assert abs(np.corrcoef(u0.cpu().numpy(), u1)[0, 1] - np.sqrt(0.75)) < 5e-3


# ---------------------------------------------------------------- importance sampler
def test_importance_discrete_vs_oracle_and_golden(cuda_dev, golden):
    import oracle
    from pmmh_qn_b200 import kernels as K
    g = golden["re_kernels"]
    for (n, nobs, seeds) in gi.RE_KERNEL_CASES:
        for seed in seeds:
            obs, params, rvr, rvp = gi.re_inputs(n, nobs, seed)
            out = K.importance_discrete(_t(obs, cuda_dev), _t(params.reshape(1, 2), cuda_dev),
                                        _t(np.array([rvr]), cuda_dev), _t(rvp.reshape(1, -1), cuda_dev),
                                        nobs, n)
            ref = oracle.importance_discrete(obs, params, rvr, rvp, n)
            pre = "is_n%d_t%d_s%d_" % (n, nobs, seed)
            ll = float(out["log_like"][0])
            assert abs(ll - ref["log_like"]) <= 1e-12 * abs(ref["log_like"])
            assert abs(ll - float(g[pre + "log_like"])) <= 1e-12 * abs(ll)
            assert int(out["traj_idx"][0]) == ref["traj_idx"]
            assert relerr(out["filt"][0].cpu().numpy(), g[pre + "filt"]) <= 1e-12
            assert relerr(out["traj"][0].cpu().numpy(), g[pre + "traj"]) <= 1e-14
            gr = out["gradient"][0].cpu().numpy()
            assert np.max(np.abs(gr - g[pre + "gradient"])) <= 1e-10 * np.max(np.abs(g[pre + "gradient"]))


def test_importance_discrete_batch(cuda_dev):
    import oracle
    from pmmh_qn_b200 import kernels as K
    n, nobs, B = 100, 100, 257
    obs = gi.re_obs(nobs)
    rs = np.random.RandomState(11)
    params = np.stack([1.0 + 0.1 * rs.normal(size=B), 0.2 + 0.02 * rs.uniform(size=B)], axis=1)
    rvs = rs.normal(size=(B, nobs, n + 1))
    rvr = norm.cdf(rvs[:, 0, 0])
    rvp = np.ascontiguousarray(rvs[:, :, 1:]).reshape(B, -1)
    out = K.importance_discrete(_t(obs, cuda_dev), _t(params, cuda_dev), _t(rvr, cuda_dev),
                                _t(rvp, cuda_dev), nobs, n)
    ll = out["log_like"].cpu().numpy()
    gr = out["gradient"].cpu().numpy()
    for b in (0, 1, 100, 256):
        ref = oracle.importance_discrete(obs, params[b], rvr[b], rvp[b], n)
        assert abs(ll[b] - ref["log_like"]) <= 1e-12 * abs(ref["log_like"])
        assert np.max(np.abs(gr[b] - ref["gradient"])) <= 1e-10 * np.max(np.abs(ref["gradient"]))


# ---------------------------------------------------------------- subsampling
@pytest.mark.parametrize("m,n_data", [(5500, 110000), (77, 1000), (550000, 11000000), (1, 5), (64, 64)])
def test_subsample_indices_exact(cuda_dev, golden, m, n_data):
    """sort(Phi(u)) + stratified must equal the reference's merge walk index for index."""
    import oracle
    from pmmh_qn_b200 import kernels as K
    for seed in (0, 1):
        u = gi.logit_u(m, seed)
        r = np.sort(norm.cdf(u))
        want = oracle.stratified(r, n_data)
        key = "strat_m%d_n%d_s%d" % (m, n_data, seed)
        if key in golden["ss_kernels"].files:
            assert np.array_equal(want, golden["ss_kernels"][key])
        # (a) host-computed Phi -> device sort + closed-form stratified: bit-exact path
        idx, srt = K.subsample_indices(_t(norm.cdf(u), cuda_dev), n_data, apply_cdf=False,
                                       want_sorted=True)
        assert np.array_equal(srt.cpu().numpy(), r)
        assert np.array_equal(idx.cpu().numpy(), want)
        # (b) device Phi: uniforms may differ in the last bits; indices may move by one at
        #     cut-point ties only
        idx2 = K.subsample_indices(_t(u, cuda_dev), n_data, apply_cdf=True).cpu().numpy()
        d = np.abs(idx2.astype(np.int64) - want.astype(np.int64))
        assert d.max() <= 1 and np.count_nonzero(d) <= max(2, m // 100000)


def test_subsample_sort_with_ties_and_edges(cuda_dev):
    from pmmh_qn_b200 import kernels as K
    r = np.array([0.5] * 40 + [0.0, 1.0, 0.25, 0.25, 1.0 - 2 ** -53] + [0.75] * 19, dtype=np.float64)
    idx, srt = K.subsample_indices(_t(r, cuda_dev), 1000, apply_cdf=False, want_sorted=True)
    assert np.array_equal(srt.cpu().numpy(), np.sort(r))
    import oracle
    assert np.array_equal(idx.cpu().numpy(), oracle.stratified(np.sort(r), 1000))


@pytest.mark.parametrize("d", [22, 28, 32, 5])
@pytest.mark.parametrize("hess", [0, 1])
def test_logistic_loglike_vs_numpy_oracle(cuda_dev, d, hess):
    import oracle
    from pmmh_qn_b200 import kernels as K
    n_data, m = 50000, 2500
    x, y, beta = gi.logit_data(n_data, d, seed=d)
    idx = oracle.subsample_indices(gi.logit_u(m, 5), n_data)
    ref = oracle.logistic_loglike_gradient(beta, x, y, idx, True, bool(hess))
    out = K.logistic_loglike(_t(x, cuda_dev), _t(y, cuda_dev), _t(idx.astype(np.int32), cuda_dev),
                             _t(beta, cuda_dev), compute_hessian=bool(hess)).cpu().numpy()
    assert abs(out[0] - ref["log_like"]) <= 1e-11 * abs(ref["log_like"])
    assert np.max(np.abs(out[1:1 + d] - ref["gradient"])) <= 1e-10 * np.max(np.abs(ref["gradient"]))
    if hess:
        h = out[1 + d:].reshape(d, d)
        assert np.max(np.abs(h - ref["hessian"])) <= 1e-10 * np.max(np.abs(ref["hessian"]))
    else:
        assert np.all(out[1 + d:] == 0.0)


def test_logistic_row_shards_sum_to_whole(cuda_dev):
    """Row-sharded evaluation (the multi-GPU decomposition) adds up to the unsharded result."""
    import oracle
    from pmmh_qn_b200 import kernels as K
    n_data, m, d = 30000, 3000, 28
    x, y, beta = gi.logit_data(n_data, d, seed=2)
    idx = _t(oracle.subsample_indices(gi.logit_u(m, 9), n_data).astype(np.int32), cuda_dev)
    whole = K.logistic_loglike(_t(x, cuda_dev), _t(y, cuda_dev), idx, _t(beta, cuda_dev),
                               compute_hessian=True).cpu().numpy()
    parts = np.zeros_like(whole)
    bounds = [0, 7000, 7001, 20000, n_data]
    for lo, hi in zip(bounds[:-1], bounds[1:]):
        parts += K.logistic_loglike(_t(x[lo:hi], cuda_dev), _t(y[lo:hi], cuda_dev), idx,
                                    _t(beta, cuda_dev), compute_hessian=True, row_begin=lo,
                                    row_end=hi).cpu().numpy()
    assert np.max(np.abs(parts - whole)) <= 1e-11 * np.max(np.abs(whole))
