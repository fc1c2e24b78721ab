"""Lock-step sampler front-end on the GPU: B chains of a sampler with the reference's interface, their
smoother calls batched into one launch (parameter/lockstep.py, state/particle_methods/batched.py).
Batching is transparent: every chain equals the same chain run alone over ParticleMethodsCUDA, bit for
bit, with host u (NumPy arrays, the reference's Crank-Nicolson) and with u resident on the device."""
import copy

import numpy as np
import pytest

import golden_inputs as gi
from toy_models import ToySVModel

pytestmark = pytest.mark.gpu


def _sampler(nobs, seed, **kw):
    from pmmh_qn_b200.parameter.cpmh import CorrelatedPMMH
    model = ToySVModel(gi.sv_obs(nobs), gi.SV_PARAM_SETS[0])
    st = {'no_iters': 7, 'no_burnin_iters': 2, 'initial_params': (0.2 + 0.01 * seed, 0.9, 0.4, -0.5), 'step_size': 0.02,
          'precond': (1.0, 0.01, 0.05, 0.05), 'drift': True, 'correlated_rvs_sigma': 0.4}
    st.update(kw)
    return CorrelatedPMMH(model, st)


def _trace(s):
    h = s.state_history
    return (np.array([h[i]['params'] for i in range(len(h))]), np.array([h[i]['log_like'] for i in range(len(h))]),
            np.array([h[i]['accepted'] for i in range(len(h))]))


def test_chains_in_lockstep_equal_chains_alone_host_rvs(cuda_dev):
    from pmmh_qn_b200 import ParticleMethodsCUDA
    from pmmh_qn_b200.parameter.lockstep import LockstepRunner
    from pmmh_qn_b200.state.particle_methods.batched import BatchedParticleMethodsCUDA
    n, nobs, B = 500, 61, 5
    seeds = [100 + k for k in range(B)]
    solo = []
    for k in range(B):
        s = _sampler(nobs, k)
        est = ParticleMethodsCUDA(s.model, no_particles=n)
        np.random.seed(seeds[k])
        s.run(est)
        solo.append(_trace(s))
    chains = [_sampler(nobs, k) for k in range(B)]
    backend = BatchedParticleMethodsCUDA(chains[0].model, no_particles=n)
    runner = LockstepRunner(chains, backend, seeds=seeds).run()
    assert runner.no_batches == 7 and set(runner.batch_sizes) == {B} and backend.no_launches == 7
    for k in range(B):
        got = _trace(chains[k])
        for a, b in zip(got, solo[k]):
            assert np.array_equal(a, b), k
    assert any(t[2][1:].sum() > 0 for t in solo) and any(t[2][1:].sum() < 6 for t in solo)   # accepts and rejects happened


def test_batched_backend_uses_device_rows_in_place(cuda_dev):
    """Handles that are the rows of one [B, NOBS, N] tensor are evaluated without a copy and give what
    ParticleMethodsCUDA gives for the same handle one by one."""
    import torch
    from pmmh_qn_b200 import ParticleMethodsCUDA
    from pmmh_qn_b200.parameter.cpmh import BatchedRVSState
    from pmmh_qn_b200.parameter.lockstep import _Request
    from pmmh_qn_b200.state.particle_methods.batched import BatchedParticleMethodsCUDA
    n, nobs, B = 4096, 41, 6
    state = BatchedRVSState(B, nobs, n, cuda_dev, sigma_u=0.3, seed=5)
    models = [ToySVModel(gi.sv_obs(nobs), np.array(gi.SV_PARAM_SETS[0]) + 0.01 * k) for k in range(B)]
    backend = BatchedParticleMethodsCUDA(models[0], no_particles=n)
    one = ParticleMethodsCUDA(models[0], no_particles=n)
    for proposed in (False, True):
        if proposed:
            for b in range(B):
                state.propose(b)
        reqs = [_Request(b, "smoother", models[b], {'rvs': {'rvs': state.handle(b, proposed)}}, dict(backend.settings))
                for b in range(B)]
        stacked = []
        orig = torch.stack
        torch.stack = lambda *a, **k: (stacked.append(1), orig(*a, **k))[1]
        try:
            backend.evaluate_batch(reqs)
        finally:
            torch.stack = orig
        assert len(stacked) == 1          # only the [B, NOBS] resampling normals are stacked, u is used in place
        for b in range(B):
            assert reqs[b].ok
            assert one.smoother(models[b], rvs={'rvs': state.handle(b, proposed)})
            assert reqs[b].results['log_like'] == one.results['log_like']
            assert np.array_equal(reqs[b].results['gradient_internal'], one.results['gradient_internal'])
    # accept copies the row, the current row then equals the proposal
    before = state.cur_u[2].clone()
    state.accept(2)
    assert torch.equal(state.cur_u[2], state.prop_u[2]) and not torch.equal(before, state.cur_u[2])


def test_lockstep_with_device_resident_u(cuda_dev):
    """The whole loop with u on the device: deterministic, every chain makes one batched call per iteration."""
    from pmmh_qn_b200.parameter.cpmh import BatchedRVSState, _ChainRVS
    from pmmh_qn_b200.parameter.lockstep import LockstepRunner
    from pmmh_qn_b200.state.particle_methods.batched import BatchedParticleMethodsCUDA
    n, nobs, B = 2048, 51, 4

    def run_once():
        state = BatchedRVSState(B, nobs, n, cuda_dev, sigma_u=0.3, seed=9)
        chains = [_sampler(nobs, k, rvs=_ChainRVS(state, k), no_iters=5) for k in range(B)]
        backend = BatchedParticleMethodsCUDA(chains[0].model, no_particles=n)
        runner = LockstepRunner(chains, backend, seeds=[7 + k for k in range(B)]).run()
        assert backend.no_launches == 5 and set(runner.batch_sizes) == {B}
        return [_trace(c) for c in chains]

    a, b = run_once(), run_once()
    for ta, tb in zip(a, b):
        for x, y in zip(ta, tb):
            assert np.array_equal(x, y)
    assert all(np.all(np.isfinite(t[1])) for t in a)
