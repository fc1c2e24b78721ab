"""Batched sampler front-end: B Metropolis-Hastings chains in lock-step, their estimator calls batched
into one device launch (BASELINE configs[3]: 1024 chains x N = 4096).

The chains are ordinary sampler objects with the reference's interface -- `sampler.run(estimator)` calling
`estimator.smoother(model, rvs={'rvs': ...})` / `estimator.filter(...)` and reading `estimator.results`
(`parameter/mcmc/base_class.py:76-160`, `mh_quasi_newton.py:303-418`) -- and they are NOT modified: the
reference's own `QuasiNewtonMetropolisHastings` runs under this front-end unchanged (tests/
test_lockstep_reference.py does exactly that where /root/reference exists).  Each chain runs as a
coroutine (a thread that only ever runs while it holds the baton) against a per-chain proxy estimator;
when every live chain has arrived at its next estimator call, the runner hands all pending calls to the
batched backend in ONE call (`backend.evaluate_batch(requests)`, on the GPU one launch of the chain
kernel over all chains: state/particle_methods/batched.py) and resumes the chains with their own `results`.

Chains keep their own random streams: the global NumPy generator the reference's samplers draw from is
saved and restored around every slice a chain runs, so chain k produces exactly the numbers it would
produce alone after `np.random.seed(seeds[k])` (or from a captured generator state) -- B chains in
lock-step equal B separate runs.
"""
import copy
import threading

import numpy as np


class _Request(object):
    __slots__ = ("chain", "kind", "model", "kwargs", "settings", "ok", "results", "extra")

    def __init__(self, chain, kind, model, kwargs, settings):
        self.chain, self.kind, self.model, self.kwargs, self.settings = chain, kind, model, kwargs, settings
        self.ok, self.results, self.extra = False, None, None


class ChainEstimator(object):
    """What one chain sees as its estimator: the attributes the samplers read (`alg_type`, `dim_rvs`,
    `settings`, `results`; base_class.py:100-101,169-179) and `smoother` / `filter`, which park the call
    until the whole batch has been evaluated."""

    def __init__(self, runner, index, backend):
        self._runner, self._index, self._backend = runner, index, backend
        self.alg_type = backend.alg_type
        self.dim_rvs = backend.dim_rvs
        self.settings = copy.deepcopy(backend.settings)
        self.results = {}
        self.diagnostics = {}
        self.no_calls = 0

    def __getattr__(self, name):
        # anything else a sampler reads (name, no_particles, ...) comes from the backend's estimator
        if name.startswith("_"):
            raise AttributeError(name)
        return getattr(self._backend.template, name)

    def _call(self, kind, model, kwargs):
        req = _Request(self._index, kind, model, kwargs, self.settings)
        self._runner._park(self._index, req)       # returns when the batch has been evaluated
        self.no_calls += 1
        if req.results is not None:
            self.results = req.results
        if req.extra is not None:
            self.diagnostics = req.extra
        return bool(req.ok)

    def smoother(self, model, **kwargs):
        return self._call("smoother", model, kwargs)

    def filter(self, model, **kwargs):
        return self._call("filter", model, kwargs)


class LockstepRunner(object):
    """runner = LockstepRunner(samplers, backend, seeds); runner.run(); samplers[k].state_history ..."""

    def __init__(self, samplers, backend, seeds=None, run_kwargs=None):
        self.samplers = list(samplers)
        self.backend = backend
        self.no_chains = len(self.samplers)
        self.seeds = list(seeds) if seeds is not None else list(range(self.no_chains))
        self.run_kwargs = run_kwargs or {}
        self.estimators = [ChainEstimator(self, k, backend) for k in range(self.no_chains)]
        self.errors = [None] * self.no_chains
        self.no_batches = 0
        self.batch_sizes = []
        # the baton: one semaphore per chain ("run now") and one for the runner ("the chain has parked or
        # finished"); exactly one of the B + 1 threads is ever runnable
        self._go = [threading.Semaphore(0) for _ in range(self.no_chains)]
        self._back = threading.Semaphore(0)
        self._pending = [None] * self.no_chains
        self._done = [False] * self.no_chains
        self._rng = [None] * self.no_chains

    # ---- chain side ---------------------------------------------------------------------------
    def _park(self, k, req):
        self._pending[k] = req
        self._rng[k] = np.random.get_state()
        self._back.release()
        self._go[k].acquire()
        np.random.set_state(self._rng[k])

    def _chain_main(self, k):
        self._go[k].acquire()
        seed = self.seeds[k]
        if isinstance(seed, tuple):
            np.random.set_state(seed)           # a captured generator state (e.g. after the sampler was constructed)
        else:
            np.random.seed(seed)
        try:
            self.samplers[k].run(self.estimators[k], **self.run_kwargs)
        except BaseException as exc:            # reported by run(); the other chains go on
            self.errors[k] = exc
        self._done[k] = True
        self._rng[k] = np.random.get_state()
        self._back.release()

    # ---- runner side --------------------------------------------------------------------------
    def _resume(self, k):
        """Hand the baton to chain k and wait until it parks at its next estimator call or finishes."""
        self._pending[k] = None
        self._go[k].release()
        self._back.acquire()

    def run(self):
        outer = np.random.get_state()
        old_stack = threading.stack_size()
        try:
            threading.stack_size(512 * 1024)    # B can be in the thousands
        except (ValueError, RuntimeError):
            pass
        threads = [threading.Thread(target=self._chain_main, args=(k,), daemon=True) for k in range(self.no_chains)]
        for t in threads:
            t.start()
        try:
            threading.stack_size(old_stack)
        except (ValueError, RuntimeError):
            pass
        try:
            live = list(range(self.no_chains))
            while live:
                for k in live:
                    self._resume(k)
                live = [k for k in live if not self._done[k]]
                if not live:
                    break
                reqs = [self._pending[k] for k in live]
                self.backend.evaluate_batch(reqs)
                self.no_batches += 1
                self.batch_sizes.append(len(reqs))
        finally:
            np.random.set_state(outer)
        for t in threads:
            t.join(timeout=5.0)
        failed = [(k, e) for k, e in enumerate(self.errors) if e is not None]
        if failed:
            raise RuntimeError("chains failed: " + "; ".join("%d: %r" % ke for ke in failed[:4]))
        return self


class LoopBackend(object):
    """Backend that evaluates a batch by calling one ordinary estimator per request (no batching): the
    stand-in used to check the front-end itself, and the way to run it over any estimator class."""

    def __init__(self, estimators):
        self._est = list(estimators)
        e0 = self._est[0]
        self.alg_type, self.dim_rvs, self.settings = e0.alg_type, e0.dim_rvs, e0.settings
        self.template = e0

    def evaluate_batch(self, requests):
        for req in requests:
            est = self._est[req.chain]
            est.settings.update(req.settings)
            fn = est.smoother if req.kind == "smoother" else est.filter
            req.ok = fn(req.model, **req.kwargs)
            req.results = dict(est.results)
            req.extra = dict(getattr(est, "diagnostics", {}) or {})
