"""CPU: host-side logic of the estimator mirrors (no device work)."""
import copy

import numpy as np

import golden_inputs as gi
from toy_models import ToySVModel


def test_estimate_gradient_and_hessian_matches_reference_semantics(golden):
    """base_state_inference.py:40-101 -- prior terms added in place, selection by
    params_to_estimate, Hessian diagonal corrected with a minus sign."""
    from pmmh_qn_b200.state.base_state_inference import BaseStateInference
    g = golden["estimators"]
    pre = "sv_smoother_c0_h1_"
    model = ToySVModel(gi.sv_obs(361), gi.SV_ESTIMATOR_PARAMS[0], g[pre + "prior_grad"], g[pre + "prior_hess"])
    est = BaseStateInference()
    # start from the reference's own estimates with the priors removed again
    grad_noprior = g[pre + "log_joint_gradient_estimate"] - g[pre + "prior_grad"]
    hess_noprior = g[pre + "log_joint_hessian_estimate"] + np.diag(g[pre + "prior_hess"])
    est.results = {'log_joint_gradient_estimate': grad_noprior.copy(),
                   'log_joint_hessian_estimate': hess_noprior.copy()}
    assert est._estimate_gradient_and_hessian(model) is True
    r = est.results
    assert np.allclose(r['gradient_internal'], g[pre + "gradient_internal"], rtol=1e-13, atol=0)
    assert np.allclose(r['hessian_internal'], g[pre + "hessian_internal"], rtol=1e-12, atol=1e-12)
    assert np.allclose(r['hessian_internal_noprior'], g[pre + "hessian_internal_noprior"], rtol=1e-12, atol=1e-12)
    assert list(r['gradient'].keys()) == ['mu', 'phi', 'sigma_v', 'rho']
    est.results = {'log_like': 1.0}
    assert est._estimate_gradient_and_hessian(model) is True and 'gradient_internal' not in est.results


def test_partial_params_to_estimate():
    from pmmh_qn_b200.state.base_state_inference import BaseStateInference
    model = ToySVModel(gi.sv_obs(50), gi.SV_PARAM_SETS[0], prior_grad=[1, 2, 3, 4], prior_hess=[0.5] * 4)
    model.params_to_estimate = ('phi', 'rho')
    model.params_to_estimate_idx = np.array([1, 3])
    est = BaseStateInference()
    est.results = {'log_joint_gradient_estimate': np.array([10.0, 20.0, 30.0, 40.0]),
                   'log_joint_hessian_estimate': np.arange(16.0).reshape(4, 4)}
    assert est._estimate_gradient_and_hessian(model)
    assert np.array_equal(est.results['gradient_internal'], [22.0, 44.0])
    assert np.array_equal(est.results['hessian_internal'], [[5.0 - 0.5, 7.0], [13.0, 15.0 - 0.5]])


def test_device_rvs_handle_is_never_deep_copied():
    import torch
    from pmmh_qn_b200.parameter.rvs import DeviceRVS
    h = DeviceRVS({"u": torch.zeros(3, 4, dtype=torch.float64)}, (3, 5), "particle")
    state = {'rvs': h, 'params': np.zeros(4)}
    c = copy.deepcopy(state)
    assert c['rvs'] is h and c['params'] is not state['params']
    assert h.nbytes == 96


def test_block_ranges_cover_everything():
    from pmmh_qn_b200.sharding import block_range
    for n in (0, 1, 7, 8, 1024, 11000000):
        for world in (1, 2, 3, 8):
            seen = []
            for r in range(world):
                b, e = block_range(n, r, world)
                assert 0 <= b <= e <= n
                seen.extend(range(b, e) if n < 5000 else [])
            if n < 5000:
                assert seen == list(range(n))
            assert sum(block_range(n, r, world)[1] - block_range(n, r, world)[0] for r in range(world)) == n


def test_interval_exchange_counts_cover_every_position_once():
    from pmmh_qn_b200 import sharding as S
    rng = np.random.RandomState(3)
    for world in (1, 2, 3, 8):
        for _ in range(20):
            n_src = rng.multinomial(1000, np.ones(world) / world)
            n_dst = rng.multinomial(1000, rng.dirichlet(np.ones(world)))
            plans = [S.interval_exchange_counts(n_src, n_dst, r) for r in range(world)]
            for r in range(world):
                assert sum(plans[r][0]) == n_src[r] and sum(plans[r][1]) == n_dst[r]
                for d in range(world):
                    assert plans[r][0][d] == plans[d][1][r]


def test_split_default_capacities():
    from pmmh_qn_b200.state.particle_methods.split import default_capacities
    assert default_capacities(1 << 20, 1) == (1 << 20, 1 << 20)
    for world in (2, 4, 8):
        cap, capc = default_capacities(1 << 24, world)
        assert (1 << 24) // world < cap <= (1 << 24) and capc == (1 << 24)


def test_stream_copy_schedule_invariants():
    """Copy schedule of the host-streamed grid path (pmmh_sv_stream_schedule, no device needed): the pieces cover the
    series exactly, none crosses a slot of 256 time steps, every piece but the last ends on a multiple of 4 steps (the
    kernel reads 32-byte sectors of the staged rows), whole slots first and shorter pieces towards the end."""
    import ctypes
    from pmmh_qn_b200 import _lib
    lib = _lib.load()
    for n_obs in list(range(1, 40)) + [71, 101, 255, 256, 257, 300, 511, 512, 513, 1000, 1001, 1002, 1003, 2000, 5000]:
        buf = (ctypes.c_int * 256)()
        k = lib.pmmh_sv_stream_schedule(n_obs, buf, 256)
        assert 1 <= k <= 256
        pieces = list(buf[:k])
        assert sum(pieces) == n_obs and min(pieces) >= 1
        t0 = 0
        for i, w in enumerate(pieces):
            assert t0 // 256 == (t0 + w - 1) // 256, (n_obs, pieces)      # inside one slot
            assert (t0 + w) % 4 == 0 or i == k - 1, (n_obs, pieces)       # sector aligned
            t0 += w
    buf = (ctypes.c_int * 16)()
    assert lib.pmmh_sv_stream_schedule(1001, buf, 16) == 7
    assert list(buf[:7]) == [256, 256, 180, 76, 104, 72, 57]
    assert lib.pmmh_sv_stream_schedule(0, buf, 16) < 0
