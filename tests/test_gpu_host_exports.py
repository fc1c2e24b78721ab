"""The four host-pointer exports of the C ABI -- the entry points that mirror the reference's Cython
signatures one to one and that INTEGRATION.md's Option B binds -- called with plain NumPy arrays in
the reference's own layouts on the golden inputs (outputs of the reference's compiled kernels,
tests/golden/make_golden.py):

  pmmh_flps_sv_corr_host        <- flps_sv_corr    stochastic_volatility.pyx:205-655
  pmmh_bpf_sv_corr_host         <- bpf_sv_corr     stochastic_volatility.pyx:59-204
  pmmh_importance_discrete_host <- importance_discrete  random_effects.pyx:21-104
  pmmh_stratified_host          <- stratified      subsampling.pyx:34-51
"""
import ctypes

import numpy as np
import pytest

import golden_inputs as gi
from helpers import relerr

pytestmark = pytest.mark.gpu


def _p(a):
    return ctypes.c_void_p(a.ctypes.data)


@pytest.fixture(scope="module")
def lib(cuda_dev):
    from pmmh_qn_b200 import _lib
    return _lib.load()


def test_flps_sv_corr_host_vs_golden(lib, golden):
    g = golden["sv_kernels"]
    for (n, nobs, lag, seeds) in gi.SV_KERNEL_CASES:
        for seed in seeds:
            obs, params, rvr, rvp = gi.sv_inputs(n, nobs, seed)
            for hess in (0, 1):
                filt, smo, traj = np.empty(nobs), np.empty(nobs), np.empty(nobs)
                grad, ll = np.empty((4, nobs)), np.empty(1)
                h1, h2 = np.empty((4, 4)), np.empty((4, 4))
                diag = np.zeros(16, dtype=np.int64)
                rc = lib.pmmh_flps_sv_corr_host(_p(obs), _p(params), _p(rvr), _p(rvp), nobs, n, lag, hess,
                                                _p(filt), _p(smo), _p(ll), _p(grad), _p(traj), _p(h1), _p(h2),
                                                _p(diag))
                assert rc == 0, lib.pmmh_last_error()
                pre = "flps_n%d_t%d_l%d_s%d_h%d_" % (n, nobs, lag, seed, hess)
                ref_ll = float(g[pre + "log_like"])
                assert abs(ll[0] - ref_ll) <= 1e-10 * abs(ref_ll), pre
                assert relerr(filt, g[pre + "filt"]) <= 1e-10, pre
                assert relerr(smo, g[pre + "smo"]) <= 1e-10, pre
                assert relerr(traj, g[pre + "traj"]) <= 1e-12, pre
                gref = g[pre + "gradient"].reshape(4, nobs)
                assert np.max(np.abs(grad - gref)) <= 1e-9 * np.max(np.abs(gref)), pre
                if hess:
                    for got, key in ((h1, "hess1"), (h2, "hess2")):
                        href = g[pre + key].reshape(4, 4)
                        assert np.max(np.abs(got - href)) <= 1e-8 * np.max(np.abs(href)), pre + key
                assert diag[2] == 0


def test_flps_sv_corr_host_reports_errors(lib):
    obs, params, rvr, rvp = gi.sv_inputs(75, 361, 0)
    out = np.empty(4 * 361)
    rc = lib.pmmh_flps_sv_corr_host(_p(obs), _p(params), _p(rvr), _p(rvp), 361, 75, 1, 0, _p(out), _p(out), _p(out),
                                    _p(out), _p(out), _p(out), _p(out), None)
    assert rc != 0 and b"lag" in lib.pmmh_last_error()


def test_bpf_sv_corr_host_vs_golden(lib, golden):
    g = golden["sv_kernels"]
    for (n, nobs, lag, seeds) in gi.SV_KERNEL_CASES:
        if n > 1024:
            continue
        for seed in seeds:
            obs, params, rvr, rvp = gi.sv_inputs(n, nobs, seed)
            filt, traj, ll = np.empty(nobs), np.empty(nobs), np.empty(1)
            diag = np.zeros(16, dtype=np.int64)
            rc = lib.pmmh_bpf_sv_corr_host(_p(obs), _p(params), _p(rvr), _p(rvp), nobs, n, 0, _p(filt), _p(ll),
                                           _p(traj), _p(diag))
            assert rc == 0, lib.pmmh_last_error()
            pre = "bpf_n%d_t%d_l%d_s%d_" % (n, nobs, lag, seed)
            ref_ll = float(g[pre + "log_like"])
            if not np.isfinite(ref_ll):      # Q2: the parity read mode degenerates exactly like the reference
                assert not np.isfinite(ll[0])
                continue
            assert abs(ll[0] - ref_ll) <= 1e-10 * abs(ref_ll), pre
            assert relerr(filt, g[pre + "filt"]) <= 1e-10, pre


def test_importance_discrete_host_vs_golden(lib, golden):
    g = golden["re_kernels"]
    for (n, nobs, seeds) in gi.RE_KERNEL_CASES:
        for seed in seeds:
            obs, params, rvr, rvp = gi.re_inputs(n, nobs, seed)
            filt, traj, ll, grad = np.empty(nobs), np.empty(nobs), np.empty(1), np.empty(2 * nobs)
            rc = lib.pmmh_importance_discrete_host(_p(obs), _p(params), ctypes.c_double(rvr), _p(rvp), nobs, n,
                                                   _p(filt), _p(ll), _p(traj), _p(grad))
            assert rc == 0, lib.pmmh_last_error()
            pre = "is_n%d_t%d_s%d_" % (n, nobs, seed)
            ref_ll = float(g[pre + "log_like"])
            assert abs(ll[0] - ref_ll) <= 1e-12 * abs(ref_ll), pre
            assert relerr(filt, g[pre + "filt"]) <= 1e-12, pre
            assert relerr(traj, g[pre + "traj"]) <= 1e-14, pre
            gref = np.asarray(g[pre + "gradient"]).reshape(-1)
            assert np.max(np.abs(grad[:gref.size] - gref)) <= 1e-10 * np.max(np.abs(gref)), pre


def test_stratified_host_vs_golden(lib, golden):
    g = golden["ss_kernels"]
    for (m, n_data, seeds) in gi.SS_KERNEL_CASES:
        for seed in seeds:
            r = gi.ss_inputs(m, seed)
            idx = np.empty(m, dtype=np.int32)
            rc = lib.pmmh_stratified_host(_p(r), m, n_data, _p(idx))
            assert rc == 0, lib.pmmh_last_error()
            assert np.array_equal(idx, g["strat_m%d_n%d_s%d" % (m, n_data, seed)])
    # edge cases of subsampling.pyx:34-51: one draw, as many draws as rows
    for m, n_data in ((1, 5), (64, 64)):
        r = gi.ss_inputs(m, 0)
        idx = np.empty(m, dtype=np.int32)
        assert lib.pmmh_stratified_host(_p(r), m, n_data, _p(idx)) == 0
        import oracle
        assert np.array_equal(idx, oracle.stratified(r, n_data))
