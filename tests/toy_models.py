"""Minimal stand-ins for the reference's model objects (models/*.py) with the attributes the
estimators read.  The reference itself is not available on the GPU box; prior derivatives are
injected from the golden files where a test needs the reference's exact numbers."""
import numpy as np


class ToySVModel(object):
    short_name = 'sv'

    def __init__(self, obs, params, prior_grad=None, prior_hess=None):
        self.obs = np.asarray(obs, dtype=np.float64).reshape(-1, 1)
        self.no_obs = self.obs.shape[0] - 1
        self.params = {'mu': float(params[0]), 'phi': float(params[1]),
                       'sigma_v': float(params[2]), 'rho': float(params[3])}
        self.no_params = 4
        self.params_to_estimate = ('mu', 'phi', 'sigma_v', 'rho')
        self.params_to_estimate_idx = np.arange(4)
        self.no_params_to_estimate = 4
        self.using_gradients = True
        self.using_hessians = False
        self._pg = np.zeros(4) if prior_grad is None else np.asarray(prior_grad, dtype=np.float64)
        self._ph = np.zeros(4) if prior_hess is None else np.asarray(prior_hess, dtype=np.float64)

    def get_all_params(self):
        return np.array([self.params[k] for k in self.params])

    def log_prior_gradient(self):
        return {k: self._pg[i] for i, k in enumerate(self.params)}

    def log_prior_hessian(self):
        return {k: self._ph[i] for i, k in enumerate(self.params)}


class ToyREModel(object):
    short_name = 'random_effects'

    def __init__(self, obs, params, prior_grad=None, prior_hess=None):
        self.obs = np.asarray(obs, dtype=np.float64)
        self.no_obs = self.obs.shape[0]
        self.params = {'mu': float(params[0]), 'sigma': float(params[1])}
        self.no_params = 2
        self.params_to_estimate = ('mu', 'sigma')
        self.params_to_estimate_idx = np.arange(2)
        self.no_params_to_estimate = 2
        self.using_gradients = True
        self.using_hessians = False
        self._pg = np.zeros(2) if prior_grad is None else np.asarray(prior_grad, dtype=np.float64)
        self._ph = np.zeros(2) if prior_hess is None else np.asarray(prior_hess, dtype=np.float64)

    def get_all_params(self):
        return np.array([self.params[k] for k in self.params])

    def log_prior_gradient(self):
        return {k: self._pg[i] for i, k in enumerate(self.params)}

    def log_prior_hessian(self):
        return {k: self._ph[i] for i, k in enumerate(self.params)}


class ToyLogisticModel(object):
    short_name = 'logistic'

    def __init__(self, x, y, beta, prior_grad=None, prior_hess=None):
        self.regressors = np.asarray(x, dtype=np.float64)
        self.obs = np.asarray(y, dtype=np.float64).flatten()
        self.no_obs = self.regressors.shape[0]
        self.no_params = self.no_regressors = self.regressors.shape[1]
        self.params = np.array(beta, dtype=np.float64)
        self.params_to_estimate = range(self.no_params)
        self.params_to_estimate_idx = np.arange(self.no_params).astype(int)
        self.no_params_to_estimate = self.no_params
        self.using_gradients = True
        self.using_hessians = False
        self._pg = np.zeros(self.no_params) if prior_grad is None else np.asarray(prior_grad)
        self._ph = np.zeros(self.no_params) if prior_hess is None else np.asarray(prior_hess)

    def get_all_params(self):
        return np.array(self.params)

    def log_prior_gradient(self):
        return np.array(self._pg, copy=True)

    def log_prior_hessian(self):
        return np.array(self._ph, copy=True)
