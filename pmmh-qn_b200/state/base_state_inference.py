"""Base class of the CUDA estimators -- the drop-in boundary.

Mirrors the contract of the reference's ``BaseStateInference``
(/root/reference/python/state/base_state_inference.py:24-101): estimators expose
``filter(model, **kw) -> bool``, ``smoother(model, **kw) -> bool``, ``results``, ``settings``,
``dim_rvs``, ``alg_type`` and add the log-prior derivatives of the (Python) model object to
the estimated gradient / Hessian.  Samplers of the reference (parameter/mcmc/*.py) can call
these classes unchanged.

Unlike the reference, importing this module does NOT switch ``warnings`` to errors globally
(base_state_inference.py:22); non-finite device results are turned into ``False`` explicitly.
"""
import numpy as np


class BaseStateInference(object):
    name = []
    settings = {}
    results = {}
    model = {}

    no_obs = 0
    log_like = []
    gradient = []
    gradient_internal = []
    hessian_internal = []

    def __repr__(self):
        return str(self.name)

    def _estimate_gradient_and_hessian(self, model):
        """Adds the gradient / Hessian of the log-prior to the likelihood estimates and
        selects the parameters under inference (base_state_inference.py:40-101)."""
        res = self.results
        have_grad = 'log_joint_gradient_estimate' in res
        have_hess = 'log_joint_hessian_estimate' in res
        if have_hess:
            hess_est = res['log_joint_hessian_estimate']
            idx = model.params_to_estimate_idx
            res['hessian_internal_noprior'] = np.copy(hess_est[np.ix_(idx, idx)])
        if not have_grad and not have_hess:
            return True
        grad_est = res['log_joint_gradient_estimate']

        prior_grad = model.log_prior_gradient()
        gradient = {}
        gradient_internal = []
        if type(model.params) is dict:
            keys = list(model.params.keys())
        else:
            keys = list(range(model.no_params))
        for i, key in enumerate(keys):
            grad_est[i] += prior_grad[key]          # in place, as the reference does
            if key in model.params_to_estimate:
                gradient[key] = grad_est[i]
                gradient_internal.append(grad_est[i])

        if have_hess:
            prior_hess = model.log_prior_hessian()
            for i, key in enumerate(keys):
                hess_est[i, i] -= prior_hess[key]

        res['gradient_internal'] = np.array(gradient_internal)
        res['gradient'] = gradient
        if have_hess:
            idx = model.params_to_estimate_idx
            res['hessian_internal'] = np.array(hess_est[np.ix_(idx, idx)])
            res['hessian_internal_prior'] = np.array(hess_est[np.ix_(idx, idx)])
        return True
