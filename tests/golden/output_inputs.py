"""A seeded stand-in for a finished sampler (what parameter/mcmc/output.py reads from it), shared by
make_output_golden.py and tests/test_output_writer.py."""
import types

import numpy as np


def fake_sampler(benchmark=False, seed=5):
    rs = np.random.RandomState(seed)
    no_iters, burn, d, nobs = 12, 4, 3, 9
    hist = {}
    for i in range(no_iters):
        st = {'params': rs.normal(size=d), 'params_prop': rs.normal(size=d), 'nat_gradient': rs.normal(size=d),
              'accepted': float(i % 2), 'state_trajectory': rs.normal(size=nobs + 1)}
        if benchmark:
            st.update({'error_bfgs_fro': float(rs.uniform()), 'error_ls_fro': float(rs.uniform()),
                       'error_sr1_fro': float(rs.uniform())})
        hist[i] = st
    hist[5]['nat_gradient'][1] = np.nan
    hist[6]['params_prop'][0] = np.inf
    hist[7]['state_trajectory'][2] = -np.inf
    hist[8]['accepted'] = np.nan
    model = types.SimpleNamespace(obs=rs.normal(size=(nobs + 1, 1)), states=None)
    settings = {'no_iters': no_iters, 'no_burnin_iters': burn, 'step_size': 0.5, 'hessian': np.eye(2),
                'correlated_rvs': True, 'initial_params': (1.0, 2.0, 3.0)}
    smp = types.SimpleNamespace(settings=settings, state_history=hist, time_per_iter=0.0125, model=model,
                                name='qmh_bfgs', adapted_step_sizes=[0.1, 0.2], no_hessians_corrected=3)
    return smp
