"""Drop-in level: the CUDA estimator classes against results recorded from the reference's own
Python estimator classes (tests/golden/estimators.npz) and from a run of its unmodified
quasi-Newton sampler (tests/golden/qn_chain.npz)."""
import copy

import numpy as np
import pytest

import golden_inputs as gi
from helpers import relerr
from toy_models import ToyLogisticModel, ToyREModel, ToySVModel

pytestmark = pytest.mark.gpu


def test_particle_methods_cuda_vs_reference_estimator(cuda_dev, golden):
    from pmmh_qn_b200 import ParticleMethodsCUDA
    g = golden["estimators"]
    n, nobs = 75, 361
    for ci, params in enumerate(gi.SV_ESTIMATOR_PARAMS):
        for hess in (0, 1):
            pre = "sv_smoother_c%d_h%d_" % (ci, hess)
            model = ToySVModel(gi.sv_obs(nobs), params, g[pre + "prior_grad"], g[pre + "prior_hess"])
            model.using_hessians = bool(hess)
            est = ParticleMethodsCUDA(model, no_particles=n, fixed_lag=10)
            assert est.dim_rvs == (nobs, n + 1) and est.alg_type == 'particle'
            assert est.settings['no_particles'] == n and est.settings['no_obs'] == nobs
            ok = est.smoother(model, rvs={'rvs': gi.sv_rvs(n, nobs, seed=1000 + ci)})
            assert ok == bool(g[pre + "ok"])
            r = est.results
            assert abs(r['log_like'] - float(g[pre + "log_like"])) <= 1e-10 * abs(r['log_like'])
            for k in ('filt_state_est', 'state_trajectory', 'smo_state_est'):
                assert relerr(r[k], g[pre + k]) <= 1e-10, k
            gi_ = g[pre + "gradient_internal"]
            assert np.max(np.abs(r['gradient_internal'] - gi_)) <= 1e-9 * np.max(np.abs(gi_))
            assert set(r['gradient'].keys()) == {'mu', 'phi', 'sigma_v', 'rho'}
            if hess:
                for k in ('hessian_internal', 'hessian_internal_noprior', 'log_joint_hessian_estimate'):
                    assert np.max(np.abs(r[k] - g[pre + k])) <= 1e-8 * np.max(np.abs(g[pre + k])), k
        pre = "sv_filter_c%d_" % ci
        model = ToySVModel(gi.sv_obs(nobs), params)
        est = ParticleMethodsCUDA(model, no_particles=n)
        assert est.filter(model, rvs={'rvs': gi.sv_rvs(n, nobs, seed=1000 + ci)})
        assert abs(est.results['log_like'] - float(g[pre + "log_like"])) <= 1e-10 * abs(float(g[pre + "log_like"]))
        assert relerr(est.results['filt_state_est'], g[pre + "filt_state_est"]) <= 1e-10
        assert relerr(est.results['state_trajectory'], g[pre + "state_trajectory"]) <= 1e-12


def test_device_rvs_handle_equals_numpy_path(cuda_dev):
    """A device-resident u handle gives the same estimate as the NumPy array it came from,
    survives copy.deepcopy (mh_quasi_newton.py:232) without copying, and the device
    Crank-Nicolson proposal matches the host formula."""
    import oracle
    from pmmh_qn_b200 import DeviceRVS, ParticleMethodsCUDA, propose_rvs
    n, nobs = 500, 200
    model = ToySVModel(gi.sv_obs(nobs), gi.SV_PARAM_SETS[0])
    est = ParticleMethodsCUDA(model, no_particles=n)
    rvs = gi.sv_rvs(n, nobs, 3)
    assert est.smoother(model, rvs={'rvs': rvs})
    ll_np, g_np = est.results['log_like'], est.results['gradient_internal'].copy()
    h = DeviceRVS.from_numpy_particle(rvs, cuda_dev)
    assert copy.deepcopy({'rvs': h})['rvs'] is h
    assert np.array_equal(h.to_numpy_particle(), rvs)
    assert est.smoother(model, rvs={'rvs': h})
    assert abs(est.results['log_like'] - ll_np) <= 1e-10 * abs(ll_np)
    assert np.max(np.abs(est.results['gradient_internal'] - g_np)) <= 1e-9 * np.max(np.abs(g_np))
    xi = gi.sv_rvs(n, nobs, 4)
    prop = propose_rvs(h, 0.5, xi=DeviceRVS.from_numpy_particle(xi, cuda_dev))
    assert np.array_equal(prop.to_numpy_particle(), oracle.crank_nicolson(rvs, xi, 0.5))
    prop2 = propose_rvs(h, 0.5, seed=9)
    assert prop2.tensors['u'].shape == h.tensors['u'].shape and prop2 is not h


def test_correlated_rvs_state_two_slots_and_stale_handles(cuda_dev):
    """Device-resident CPMH state: the proposal is written into the spare slot (no allocation), accept
    swaps the slots, reject keeps the current one; a history handle whose slot has been written again
    is stale and refuses to be read (base_class.py:221-241, :269-300; mh_quasi_newton.py:226-233)."""
    import oracle
    from pmmh_qn_b200 import CorrelatedRVSState, DeviceRVS, ParticleMethodsCUDA
    n, nobs, sigma_u = 400, 120, 0.3
    model = ToySVModel(gi.sv_obs(nobs), gi.SV_PARAM_SETS[0])
    est = ParticleMethodsCUDA(model, no_particles=n)
    rvs0 = gi.sv_rvs(n, nobs, 5)
    st = CorrelatedRVSState.from_numpy_particle(rvs0, cuda_dev, sigma_u, seed=3)
    cur0 = st.current
    hist = copy.deepcopy({'rvs': cur0})                      # what the sampler's history does
    assert hist['rvs'] is cur0 and not cur0.stale
    ptrs = sorted(t.data_ptr() for slot in st._slots for t in slot.values())
    # proposal with supplied noise = the reference's formula bit for bit; reject keeps the current state
    xi = gi.sv_rvs(n, nobs, 6)
    prop = st.propose(xi=DeviceRVS.from_numpy_particle(xi, cuda_dev))
    want = oracle.crank_nicolson(rvs0, xi, sigma_u)
    assert np.array_equal(prop.to_numpy_particle(), want)
    st.reject()
    assert np.array_equal(st.current.to_numpy_particle(), rvs0) and not cur0.stale
    # accept swaps; the estimator sees the accepted u
    prop = st.propose(xi=DeviceRVS.from_numpy_particle(xi, cuda_dev))
    assert est.smoother(model, rvs={'rvs': prop})
    ll_prop = est.results['log_like']
    st.accept()
    assert np.array_equal(st.current.to_numpy_particle(), want) and not cur0.stale   # old slot still intact
    assert est.smoother(model, rvs={'rvs': want})
    assert abs(est.results['log_like'] - ll_prop) <= 1e-10 * abs(ll_prop)
    # the next proposal reuses the slot cur0 views: that handle is stale from now on
    p2 = st.propose()                                        # Philox normals on the device
    assert cur0.stale
    with pytest.raises(RuntimeError):
        cur0.tensors
    assert not p2.stale and not st.current.stale
    assert sorted(t.data_ptr() for slot in st._slots for t in slot.values()) == ptrs   # never reallocated
    u_cur, u_new = st.current.tensors['u'], p2.tensors['u']
    resid = (u_new - np.sqrt(1.0 - sigma_u ** 2) * u_cur) / sigma_u         # = the Philox normals
    assert abs(float(resid.mean())) < 0.02 and abs(float(resid.std()) - 1.0) < 0.02
    assert st.proposals == 3 and st.accepted == 1


def test_importance_sampling_cuda_vs_reference_estimator(cuda_dev, golden):
    from pmmh_qn_b200 import ImportanceSamplingCUDA
    g = golden["estimators"]
    n, nobs = 100, 100
    for ci, params in enumerate(gi.RE_ESTIMATOR_PARAMS):
        pre = "re_smoother_c%d_" % ci
        model = ToyREModel(gi.re_obs(nobs), params, g[pre + "prior_grad"], g[pre + "prior_hess"])
        est = ImportanceSamplingCUDA(model, no_particles=n)
        assert est.dim_rvs == (nobs, n + 1) and est.alg_type == 'particle'
        ok = est.smoother(model, rvs={'rvs': gi.re_rvs(n, nobs, seed=2000 + ci)})
        assert ok == bool(g[pre + "ok"])
        r = est.results
        assert abs(r['log_like'] - float(g[pre + "log_like"])) <= 1e-12 * abs(r['log_like'])
        assert relerr(r['filt_state_est'], g[pre + "filt_state_est"]) <= 1e-12
        assert relerr(r['state_trajectory'], g[pre + "state_trajectory"]) <= 1e-14
        gi_ = g[pre + "gradient_internal"]
        assert np.max(np.abs(r['gradient_internal'] - gi_)) <= 1e-10 * np.max(np.abs(gi_))


def test_direct_computation_cuda_vs_reference_estimator(cuda_dev, golden):
    from pmmh_qn_b200 import DirectComputationCUDA
    g = golden["estimators"]
    n_data, d, m = 110000, gi.LOGIT_D, 5500
    x, y, beta = gi.logit_data(n_data, d)
    pre0 = "logit_smoother_c0_h0_"
    model = ToyLogisticModel(x, y, beta, g[pre0 + "prior_grad"], g[pre0 + "prior_hess"])
    est = DirectComputationCUDA(model, no_particles=m)
    assert est.dim_rvs == m and est.alg_type == 'direct'
    for ci in range(2):
        u = gi.logit_u(m, seed=3000 + ci)
        for hess in (0, 1):
            pre = "logit_smoother_c%d_h%d_" % (ci, hess)
            assert est.smoother(model, compute_hessian=bool(hess), rvs={'rvs': u})
            r = est.results
            assert abs(r['log_like'] - float(g[pre + "log_like"])) <= 1e-10 * abs(r['log_like'])
            for k in ('gradient', 'gradient_internal'):
                assert np.max(np.abs(r[k] - g[pre + k])) <= 1e-9 * np.max(np.abs(g[pre + k])), k
            if hess:
                for k in ('hessian', 'hessian_internal'):
                    assert np.max(np.abs(r[k] - g[pre + k])) <= 1e-9 * np.max(np.abs(g[pre + k])), k
        assert est.filter(model, rvs={'rvs': u})
        want = float(g["logit_filter_c%d_log_like" % ci])
        assert abs(est.results['log_like'] - want) <= 1e-10 * abs(want)


def test_calls_recorded_from_reference_qn_sampler(cuda_dev, golden):
    """Every estimator call the reference's QuasiNewtonMetropolisHastings made in a short run
    (params + u as it proposed them) is reproduced within tolerance."""
    from pmmh_qn_b200 import ParticleMethodsCUDA
    g = golden["qn_chain"]
    n, nobs = 75, 361
    obs = gi.sv_obs(nobs)
    est = None
    checked = 0
    for k in range(int(g["n_calls"])):
        key = "call%d_rvs" % k
        if key not in g.files:
            continue
        model = ToySVModel(obs, g["call%d_params" % k])
        est = est or ParticleMethodsCUDA(model, no_particles=n)
        ok = est.smoother(model, rvs={'rvs': g[key]})
        assert ok == bool(g["call%d_ok" % k])
        want_ll = float(g["call%d_log_like" % k])
        assert abs(est.results['log_like'] - want_ll) <= 1e-10 * abs(want_ll)
        checked += 1
    assert checked >= 4


def test_host_rvs_beyond_the_record_threshold(cuda_dev):
    """N = 2^23 with a host (NumPy) rvs array: the host-streamed entry point runs the streaming kernels
    with path storage and sizes its own workspace (pmmh_sv_streamed_workspace_bytes); same estimate as
    the device-resident handle of the same numbers."""
    from pmmh_qn_b200 import DeviceRVS, ParticleMethodsCUDA
    n, nobs = 1 << 23, 24
    model = ToySVModel(gi.sv_obs(nobs), gi.SV_PARAM_SETS[0])
    est = ParticleMethodsCUDA(model, no_particles=n)
    rs = np.random.RandomState(8)
    rvs = rs.normal(size=(nobs, n + 1))
    assert est.smoother(model, rvs={'rvs': rvs})
    ll_np, g_np = est.results['log_like'], est.results['gradient_internal'].copy()
    assert est.diagnostics['status'] == 0
    h = DeviceRVS.from_numpy_particle(rvs, cuda_dev)
    del rvs
    assert est.smoother(model, rvs={'rvs': h})
    assert abs(est.results['log_like'] - ll_np) <= 1e-10 * abs(ll_np)
    assert np.max(np.abs(est.results['gradient_internal'] - g_np)) <= 1e-9 * np.max(np.abs(g_np))
