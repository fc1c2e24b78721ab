"""Correlated importance sampling on the GPU -- drop-in for the reference's
``ImportanceSamplingCython`` (/root/reference/python/state/importance_sampling/cython.py:29-132)
for the random-effects model (random_effects.pyx:21-104).

``rvs`` is the reference's (n_obs, N+1) array: ``rv_r = Phi(rvs[0, 0])`` and
``rv_p = rvs[:, 1:].flatten()`` (cython.py:82-83).  The problem is tiny (80 KB at the shipped
100 x 100), so one CTA evaluates it; ``evaluate_batch`` runs many (params, u) pairs at once.
"""
import numpy as np
import torch
from scipy.stats import norm

from ... import kernels as K
from ...parameter.rvs import DeviceRVS
from ..base_state_inference import BaseStateInference


class ImportanceSamplingCUDA(BaseStateInference):
    """Importance sampling methods (CUDA, sm_100a)."""

    def __init__(self, model, no_particles=100, device=None, verbose=False):
        self.alg_type = 'importance'
        if model.short_name != 'random_effects':
            raise NameError("CUDA implementation for model missing.")
        if not torch.cuda.is_available():
            raise RuntimeError("ImportanceSamplingCUDA needs a CUDA device; there is no CPU fallback.")
        self.device = torch.device(device) if device is not None else torch.device('cuda', torch.cuda.current_device())
        self._init_importance_sampler(model, int(no_particles), verbose)
        self.results = {}

    def _inputs(self, model, kwargs):
        obs = np.ascontiguousarray(np.array(model.obs, dtype=np.float64).reshape(-1))
        params = np.asarray(model.get_all_params(), dtype=np.float64)
        if 'rvs' in kwargs:
            rvs = kwargs['rvs']['rvs']
            if isinstance(rvs, DeviceRVS):
                full = rvs.tensors['u'].reshape(self.no_obs, self.no_particles + 1)
                rv_r = K.norm_cdf(full[0:1, 0].contiguous())
                rv_p = full[:, 1:].contiguous().reshape(1, -1)
            else:
                rvs = np.asarray(rvs, dtype=np.float64)
                rv_r = torch.tensor([norm.cdf(rvs[:, 0][0])], dtype=torch.float64, device=self.device)
                full = torch.from_numpy(np.ascontiguousarray(rvs)).to(self.device, non_blocking=True)
                rv_p = full[:, 1:].contiguous().reshape(1, -1)
        else:
            rv_r = torch.tensor([np.random.uniform()], dtype=torch.float64, device=self.device)
            rv_p = torch.from_numpy(np.random.normal(size=(self.no_obs, self.no_particles)).flatten()
                                    ).to(self.device).reshape(1, -1)
        return (torch.from_numpy(obs).to(self.device), torch.from_numpy(params).to(self.device).reshape(1, 2),
                rv_r, rv_p)

    def _run(self, model, kwargs):
        obs, params, rv_r, rv_p = self._inputs(model, kwargs)
        out = K.importance_discrete(obs, params, rv_r, rv_p, self.no_obs, self.no_particles)
        flat = torch.cat([out['filt'][0], out['traj'][0], out['log_like'], out['gradient'][0]]).cpu().numpy()
        n = self.no_obs
        return flat[:n], float(flat[2 * n]), flat[n:2 * n], flat[2 * n + 1:2 * n + 3]

    def filter(self, model, **kwargs):
        """Importance sampling (random_effects.pyx:21-104)."""
        try:
            xf, ll, xtraj, _ = self._run(model, kwargs)
            self.results.update({'filt_state_est': np.array(xf).flatten()})
            self.results.update({'state_trajectory': np.array(xtraj).flatten()})
            self.results.update({'log_like': float(ll)})
            return True
        except Exception as e:
            print("Error in CUDA code for importance sampler filter.")
            print(e)
            return False

    def smoother(self, model, **kwargs):
        """Importance sampling with the gradient wrt (mu, log sigma)."""
        try:
            xf, ll, xtraj, grad = self._run(model, kwargs)
            self.results.update({'filt_state_est': np.array(xf).flatten()})
            self.results.update({'state_trajectory': np.array(xtraj).flatten()})
            self.results.update({'log_like': float(ll)})
            self.results.update({'log_joint_gradient_estimate': np.array(grad).flatten()})
            if self._estimate_gradient_and_hessian(model):
                return True
            return False
        except Exception as e:
            print("Error in CUDA code for importance sampler.")
            print(e)
            return False

    def evaluate_batch(self, obs, params, rvs):
        """B evaluations in one launch (new API; the reference has no batched call).
        obs [n_obs]; params [B, 2]; rvs [B, n_obs, N+1] host array.  Returns a dict of host
        arrays: log_like [B], gradient [B, 2], filt [B, n_obs], traj [B, n_obs]."""
        rvs = np.asarray(rvs, dtype=np.float64)
        B = rvs.shape[0]
        rv_r = torch.from_numpy(norm.cdf(rvs[:, 0, 0])).to(self.device)
        full = torch.from_numpy(np.ascontiguousarray(rvs)).to(self.device)
        rv_p = full[:, :, 1:].contiguous().reshape(B, -1)
        out = K.importance_discrete(torch.from_numpy(np.ascontiguousarray(obs, dtype=np.float64)).to(self.device),
                                    torch.from_numpy(np.ascontiguousarray(params, dtype=np.float64)).to(self.device),
                                    rv_r, rv_p, self.no_obs, self.no_particles)
        return {k: v.cpu().numpy() for k, v in out.items()}

    def _init_importance_sampler(self, model, no_particles, verbose):
        no_obs = model.no_obs
        self.name = "Importance sampling (CUDA) for " + model.short_name + " model"
        self.alg_type = 'particle'     # as the reference reports it (cython.py:114)
        self.settings = {'no_particles': no_particles,
                         'no_obs': no_obs,
                         'estimate_gradient': False
                         }
        self.no_obs = no_obs
        self.no_particles = no_particles
        self.dim_rvs = (no_obs, no_particles + 1)
        if verbose:
            print("CUDA importance sampling implementation for " + model.short_name + " initialised.")
