#!/usr/bin/env python
"""Run in a subprocess by tests/test_lockstep_reference.py (the reference's import arms warnings as errors
and stubs modules globally): the reference's UNMODIFIED QuasiNewtonMetropolisHastings and ParticleMethodsCython
(T = 360, N = 75) under the lock-step front-end.
  1. one chain == the estimator calls recorded from the plain reference run (tests/golden/qn_chain.npz);
  2. three chains in lock-step == the same three chains run one after the other, bit for bit.
Prints one JSON line."""
import contextlib
import io
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests", "golden")):
    sys.path.insert(0, p)
import make_golden  # noqa: E402

make_golden.import_reference_python()
from parameter.mcmc.mh_quasi_newton import QuasiNewtonMetropolisHastings  # noqa: E402
from state.particle_methods.cython import ParticleMethodsCython  # noqa: E402

from pmmh_qn_b200.parameter.lockstep import LockstepRunner, LoopBackend  # noqa: E402


def build(seed, no_iters=16):
    np.random.seed(seed)
    model = make_golden.make_ref_sv_model(361, (0.2, 0.9, 0.4, -0.5))
    est = ParticleMethodsCython(model)
    hessian_guess = np.diag((0.01, 0.01, 0.01, 0.001))
    settings = {"no_iters": no_iters, "no_burnin_iters": 8, "adapt_step_size": True,
                "adapt_step_size_initial": 0.1, "adapt_step_size_rate": 0.5,
                "adapt_step_size_target": 0.2, "initial_params": (2.0, 0.9, 0.4, -0.2),
                "no_iters_between_progress_reports": 1000, "correlated_rvs": True,
                "correlated_rvs_sigma": 0.5, "memory_length": 5,
                "accept_first_iterations": 5, "hessian": hessian_guess,
                "hess_corr_fallback": hessian_guess, "hess_corr_method": "flip"}
    sampler = QuasiNewtonMetropolisHastings(model, settings, qn_method="bfgs")
    return sampler, est, np.random.get_state()


class Recording(LoopBackend):
    def __init__(self, estimators):
        super().__init__(estimators)
        self.calls = []

    def evaluate_batch(self, requests):
        super().evaluate_batch(requests)
        for r in requests:
            self.calls.append((r.chain, np.array(r.model.get_all_params(), dtype=np.float64), bool(r.ok),
                               float(r.results.get("log_like", np.nan)),
                               np.array(r.results.get("gradient_internal", np.full(4, np.nan)), dtype=np.float64)))


def record_plain(sampler, est):
    """The plain reference run (no front-end) with its estimator calls recorded."""
    calls = []
    orig = est.smoother

    def rec(mdl, **kw):
        ok = orig(mdl, **kw)
        calls.append((0, np.array(mdl.get_all_params(), dtype=np.float64), bool(ok),
                      float(est.results.get("log_like", np.nan)),
                      np.array(est.results.get("gradient_internal", np.full(4, np.nan)), dtype=np.float64)))
        return ok

    est.smoother = rec
    try:
        sampler.run(est)
    except Exception:
        pass
    return calls


def same_calls(a, b):
    if len(a) != len(b):
        return False
    for x, y in zip(a, b):
        if not (np.array_equal(x[1], y[1]) and x[2] == y[2]):
            return False
        if x[2] and not (x[3] == y[3] and np.array_equal(x[4], y[4])):
            return False
    return True


out = {}
sink = io.StringIO()
with contextlib.redirect_stdout(sink):
    # ---- 1. one chain against the recorded plain run
    g = np.load(os.path.join(ROOT, "tests", "golden", "qn_chain.npz"))
    s, e, state = build(87655678)
    be = Recording([e])
    try:
        LockstepRunner([s], be, seeds=[state]).run()
    except RuntimeError as exc:          # (the plain run also trips over its post-run statistics on so short a chain)
        out["single_run_error"] = str(exc)[:120]
    n_calls = int(g["n_calls"])
    out["calls"] = [len(be.calls), n_calls]
    ok = len(be.calls) >= n_calls
    for k in range(min(len(be.calls), n_calls)):
        _, par, okk, ll, gr = be.calls[k]
        ok = ok and np.array_equal(par, g["call%d_params" % k]) and okk == bool(g["call%d_ok" % k])
        if okk:
            ok = ok and ll == float(g["call%d_log_like" % k]) and np.array_equal(gr, g["call%d_gradient_internal" % k])
    out["single_chain_equals_recorded_run"] = bool(ok)

    # ---- 2. three chains in lock-step against the same chains alone
    seeds = (11, 22, 33)
    solo = []
    for sd in seeds:
        s, e, state = build(sd, no_iters=12)
        np.random.set_state(state)
        solo.append(record_plain(s, e))
    built = [build(sd, no_iters=12) for sd in seeds]
    be3 = Recording([b[1] for b in built])
    runner = LockstepRunner([b[0] for b in built], be3, seeds=[b[2] for b in built])
    try:
        runner.run()
    except RuntimeError as exc:
        out["lockstep_run_error"] = str(exc)[:120]
    same = all(same_calls([c for c in be3.calls if c[0] == k], solo[k]) for k in range(3))
    out["three_chains_lockstep_equal_solo"] = bool(same)
    out["calls_per_chain"] = [len([c for c in be3.calls if c[0] == k]) for k in range(3)]
    out["batches"] = runner.no_batches
    out["batch_sizes"] = sorted(set(runner.batch_sizes))
print(json.dumps(out))
