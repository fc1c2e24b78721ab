"""A small correlated pseudo-marginal Metropolis-Hastings sampler with the reference's sampler interface
(`sampler.run(estimator)`, `state_history`, `settings`, `time_per_iter`, `name`; base_class.py:76-160),
and the batched device state of the auxiliary variables of B chains.

Why it exists: the reference's samplers (quasi-Newton, first / second order; parameter/mcmc/*.py) stay the
reference's own Python and run unchanged over the CUDA estimators and under the lock-step front-end
(parameter/lockstep.py) -- but they are not available on a GPU box without the reference checkout.  This
sampler is what `bench.py` and the GPU tests drive through the front-end there: the same loop (propose
parameters, Crank-Nicolson proposal of u, one smoother call, accept / reject, state history), a
preconditioned Langevin / random-walk parameter proposal instead of the BFGS machinery.  It is NOT a
restatement of the quasi-Newton proposal.
"""
import time

import numpy as np
import torch

from .. import kernels as K
from .rvs import DeviceRVS


class BatchedRVSState(object):
    """u of B chains in two [B, n_obs, N] tensors (current / proposed) plus the resampling normals
    [B, n_obs]: handles of the proposed rows are rows of ONE tensor, so the batched backend evaluates them
    in place; accept copies a row (33 MB at T = 1000, N = 4096), reject does nothing."""

    def __init__(self, no_chains, n_obs, n_particles, device, sigma_u, seed=0):
        g = torch.Generator(device=device)
        g.manual_seed(int(seed))
        self.B, self.n_obs, self.n = int(no_chains), int(n_obs), int(n_particles)
        self.sigma_u, self.seed = float(sigma_u), int(seed)
        self.cur_u = torch.randn((self.B, n_obs, n_particles), dtype=torch.float64, device=device, generator=g)
        self.cur_r = torch.randn((self.B, n_obs), dtype=torch.float64, device=device, generator=g)
        self.prop_u = torch.empty_like(self.cur_u)
        self.prop_r = torch.empty_like(self.cur_r)
        self._offset = 0
        self.shape = (n_obs, n_particles + 1)

    @property
    def nbytes(self):
        return 2 * (self.cur_u.numel() + self.cur_r.numel()) * 8

    def handle(self, b, proposed=False):
        u, r = (self.prop_u, self.prop_r) if proposed else (self.cur_u, self.cur_r)
        return DeviceRVS({"u": u[b], "r_raw": r[b]}, self.shape, "particle")

    def propose(self, b):
        """Crank-Nicolson proposal of chain b into its proposed row (Philox normals on the device)."""
        for cur, prop in ((self.cur_r, self.prop_r), (self.cur_u, self.prop_u)):
            K.crank_nicolson(cur[b], self.sigma_u, seed=self.seed, philox_offset=self._offset, out=prop[b])
            self._offset += (cur[b].numel() + 1) // 2
        return self.handle(b, proposed=True)

    def accept(self, b):
        self.cur_u[b].copy_(self.prop_u[b])
        self.cur_r[b].copy_(self.prop_r[b])


class _ChainRVS(object):
    """Per-chain view of a BatchedRVSState with the CorrelatedRVSState interface."""

    def __init__(self, state, b):
        self._s, self._b = state, b

    @property
    def current(self):
        return self._s.handle(self._b)

    def propose(self):
        return self._s.propose(self._b)

    def accept(self):
        self._s.accept(self._b)

    def reject(self):
        pass


class HostRVS(object):
    """u as a NumPy array with the reference's Crank-Nicolson proposal (base_class.py:231-233)."""

    def __init__(self, dim_rvs, sigma_u):
        self.sigma_u = float(sigma_u)
        self.current = np.random.normal(size=dim_rvs)
        self._prop = None

    def propose(self):
        self._prop = np.sqrt(1.0 - self.sigma_u ** 2) * self.current + self.sigma_u * np.random.normal(size=self.current.shape)
        return self._prop

    def accept(self):
        self.current = self._prop

    def reject(self):
        pass


class CorrelatedPMMH(object):
    """settings: no_iters, no_burnin_iters, initial_params, step_size (scalar), precond (vector, default 1),
    drift (bool: Langevin drift from `gradient_internal`), correlated_rvs_sigma, rvs (None = HostRVS, or an
    object with current / propose() / accept() / reject(), e.g. CorrelatedRVSState or a BatchedRVSState row)."""

    def __init__(self, model, settings):
        self.model = model
        self.settings = {'no_iters': 100, 'no_burnin_iters': 10, 'step_size': 0.05, 'precond': None, 'drift': True,
                         'correlated_rvs_sigma': 0.3, 'rvs': None}
        self.settings.update(settings)
        self.name = "Correlated pseudo-marginal MH (Langevin / random-walk proposal)"
        self.state_history = {}
        self.time_per_iter = 0.0

    def _set_params(self, theta):
        for k, v in zip(list(self.model.params.keys()), theta):
            self.model.params[k] = float(v)
        chk = getattr(self.model, "check_parameters", None)
        if chk is not None:
            return bool(chk())
        p = self.model.params
        return abs(p['phi']) < 1.0 and p['sigma_v'] > 0.0 and abs(p['rho']) < 1.0

    def _log_prior(self):
        fn = getattr(self.model, "log_prior_value", None)
        return float(fn()) if fn is not None else 0.0

    def _mean(self, theta, grad, eps, pre):
        if self.settings['drift'] and grad is not None:
            return theta + 0.5 * eps * eps * pre * grad
        return theta

    @staticmethod
    def _logq(x, mean, eps, pre):
        return float(-0.5 * np.sum((x - mean) ** 2 / (eps * eps * pre)))

    def run(self, estimator):
        st = self.settings
        no_iters, eps = int(st['no_iters']), float(st['step_size'])
        d = len(self.model.params)
        pre = np.ones(d) if st['precond'] is None else np.asarray(st['precond'], dtype=np.float64)
        estimator.settings['estimate_gradient'] = True
        estimator.settings['estimate_hessian'] = False
        self.model.using_gradients, self.model.using_hessians = True, False
        rvs = st['rvs'] if st['rvs'] is not None else HostRVS(estimator.dim_rvs, st['correlated_rvs_sigma'])
        t0 = time.time()
        theta = np.asarray(st['initial_params'], dtype=np.float64)
        if not self._set_params(theta) or not estimator.smoother(self.model, rvs={'rvs': rvs.current}):
            raise NameError("MCMC: Initialisation failed, check parameters.")
        cur = {'params': theta.copy(), 'params_prop': theta.copy(), 'accepted': 1.0,
               'log_like': estimator.results['log_like'], 'log_prior': self._log_prior(),
               'nat_gradient': np.array(estimator.results['gradient_internal'], dtype=np.float64),
               'state_trajectory': np.array(estimator.results['state_trajectory'])}
        hist = {0: dict(cur)}
        for i in range(1, no_iters):
            mean_f = self._mean(cur['params'], cur['nat_gradient'], eps, pre)
            prop = mean_f + eps * np.sqrt(pre) * np.random.normal(size=d)
            rvs_prop = rvs.propose()
            ok = self._set_params(prop) and estimator.smoother(self.model, rvs={'rvs': rvs_prop})
            accept_prob = 0.0
            new = None
            if ok:
                new = {'params': prop.copy(), 'params_prop': prop.copy(), 'accepted': 1.0,
                       'log_like': estimator.results['log_like'], 'log_prior': self._log_prior(),
                       'nat_gradient': np.array(estimator.results['gradient_internal'], dtype=np.float64),
                       'state_trajectory': np.array(estimator.results['state_trajectory'])}
                mean_b = self._mean(prop, new['nat_gradient'], eps, pre)
                log_a = (new['log_like'] + new['log_prior']) - (cur['log_like'] + cur['log_prior'])
                log_a += self._logq(cur['params'], mean_b, eps, pre) - self._logq(prop, mean_f, eps, pre)
                if np.isfinite(log_a):
                    accept_prob = float(np.exp(min(0.0, log_a)))
            if np.random.random(1) < accept_prob:
                rvs.accept()
                cur = new
                hist[i] = dict(cur)
            else:
                rvs.reject()
                hist[i] = dict(cur)
                hist[i]['accepted'] = 0.0
                hist[i]['params_prop'] = prop.copy()
            hist[i]['accept_prob'] = accept_prob
        self._set_params(cur['params'])
        self.state_history = hist
        self.time_per_iter = (time.time() - t0) / max(1, no_iters)
        return self
