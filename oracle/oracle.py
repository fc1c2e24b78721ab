"""ctypes / NumPy front end of the CPU oracle.

TEST INFRASTRUCTURE ONLY -- imported by tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline / ``--impl reference`` legs, never by the product
package ``pmmh-qn_b200``.

Parity status: PINNED against the compiled reference (oracle/_ref) and the
golden vectors under tests/golden/ (see tests/test_oracle_vs_golden.py).

Functions mirror the reference call signatures (reference paths relative to
/root/reference/python):

* ``flps_sv_corr``  -- state/particle_methods/stochastic_volatility.pyx:205-655
* ``bpf_sv_corr``   -- ...stochastic_volatility.pyx:61-201
* ``importance_discrete`` -- state/importance_sampling/random_effects.pyx:21-104
* ``stratified``    -- state/direct/subsampling.pyx:34-51
* ``split_rvs_particle`` -- state/particle_methods/cython.py:89-91
* ``split_rvs_importance`` -- state/importance_sampling/cython.py:82-83
* ``smoother_post`` -- state/particle_methods/cython.py:100-126
* ``logistic_loglike_gradient`` -- models/logistic_regression.py:108-176 (NumPy)
* ``subsample_indices`` -- state/direct/standard.py:75-76
* ``crank_nicolson`` -- parameter/mcmc/base_class.py:221-241
"""
import ctypes
import os
import subprocess

import numpy as np
from scipy.stats import norm

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "_build", "libpmmh_oracle.so")
_lib = None

_dp = ctypes.POINTER(ctypes.c_double)
_ip = ctypes.POINTER(ctypes.c_int)


def build(force=False):
    """Compile oracle/pmmh_oracle.c with gcc (plain -O2)."""
    src = os.path.join(_HERE, "pmmh_oracle.c")
    if (not force and os.path.exists(_LIB_PATH)
            and os.path.getmtime(_LIB_PATH) >= os.path.getmtime(src)):
        return _LIB_PATH
    os.makedirs(os.path.dirname(_LIB_PATH), exist_ok=True)
    subprocess.check_call(
        ["gcc", "-O2", "-fPIC", "-fno-strict-overflow", "-ffp-contract=off", "-w", "-shared",
         src, "-o", _LIB_PATH, "-lm"])
    return _LIB_PATH


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = ctypes.CDLL(_LIB_PATH)
        _lib.oracle_my_max.restype = ctypes.c_double
    return _lib


def _d(a):
    return a.ctypes.data_as(_dp)


def _i(a):
    return a.ctypes.data_as(_ip)


def _c64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def argsort(data):
    data = _c64(data)
    order = np.empty(data.shape[0], dtype=np.int32)
    lib().oracle_argsort(_d(data), _i(order), ctypes.c_int(data.shape[0]))
    return order


def systematic_corr(weights, rnd):
    weights = _c64(weights)
    n = weights.shape[0]
    anc = np.empty(n, dtype=np.int32)
    lib().oracle_systematic_corr(_i(anc), _d(weights), ctypes.c_double(rnd), ctypes.c_int(n), None)
    return anc


def my_max(w):
    w = _c64(w)
    return float(lib().oracle_my_max(_d(w), ctypes.c_int(w.shape[0])))


def flps_sv_corr(obs, params, rvr, rvp, n_particles, lag=10, compute_hessian=0, dumps=False):
    """Returns dict with the 7 reference outputs (+ X, A, W, info when dumps)."""
    obs, params, rvr, rvp = _c64(obs), _c64(params), _c64(rvr), _c64(rvp)
    nobs = obs.shape[0]
    n = int(n_particles)
    assert rvr.shape[0] >= nobs and rvp.shape[0] == nobs * n
    filt = np.zeros(nobs)
    smo = np.zeros(nobs)
    ll = np.zeros(1)
    grad = np.zeros((4, nobs))
    traj = np.zeros(nobs)
    h1 = np.zeros((4, 4))
    h2 = np.zeros((4, 4))
    info = np.zeros(4, dtype=np.int32)
    if dumps:
        X = np.zeros((nobs, n))
        A = np.zeros((nobs, n), dtype=np.int32)
        W = np.zeros((nobs, n))
        xp, ap, wp = _d(X), _i(A), _d(W)
    else:
        X = A = W = None
        xp = ap = wp = None
    rc = lib().oracle_flps_sv_corr(_d(obs), _d(params), _d(rvr), _d(rvp), ctypes.c_int(n),
                                   ctypes.c_int(nobs), ctypes.c_int(lag),
                                   ctypes.c_int(int(compute_hessian)), _d(filt), _d(smo), _d(ll),
                                   _d(grad), _d(traj), _d(h1), _d(h2), xp, ap, wp, _i(info))
    if rc != 0:
        raise MemoryError("oracle_flps_sv_corr failed")
    out = dict(filt=filt, smo=smo, log_like=float(ll[0]), gradient=grad, traj=traj, hess1=h1,
               hess2=h2, traj_idx=int(info[0]), traj_oob=int(info[1]))
    if dumps:
        out.update(X=X, A=A, W=W)
    return out


def bpf_sv_corr(obs, params, rvr, rvp, n_particles, intended_read=False, dumps=False):
    obs, params, rvr, rvp = _c64(obs), _c64(params), _c64(rvr), _c64(rvp)
    nobs = obs.shape[0]
    n = int(n_particles)
    assert rvp.shape[0] == nobs * n
    filt = np.zeros(nobs)
    ll = np.zeros(1)
    traj = np.zeros(nobs)
    info = np.zeros(4, dtype=np.int32)
    if dumps:
        X = np.zeros((nobs, n))
        A = np.zeros((nobs, n), dtype=np.int32)
        W = np.zeros((nobs, n))
        xp, ap, wp = _d(X), _i(A), _d(W)
    else:
        X = A = W = None
        xp = ap = wp = None
    rc = lib().oracle_bpf_sv_corr(_d(obs), _d(params), _d(rvr), _d(rvp), ctypes.c_int(n),
                                  ctypes.c_int(nobs), ctypes.c_int(int(intended_read)), _d(filt),
                                  _d(ll), _d(traj), xp, ap, wp, _i(info))
    if rc != 0:
        raise MemoryError("oracle_bpf_sv_corr failed")
    out = dict(filt=filt, log_like=float(ll[0]), traj=traj, traj_idx=int(info[0]),
               traj_oob=int(info[1]))
    if dumps:
        out.update(X=X, A=A, W=W)
    return out


def importance_discrete(obs, params, rvr, rvp, n_particles):
    obs, params, rvp = _c64(obs), _c64(params), _c64(rvp)
    nobs = obs.shape[0]
    n = int(n_particles)
    assert rvp.shape[0] == nobs * n
    filt = np.zeros(nobs)
    ll = np.zeros(1)
    traj = np.zeros(nobs)
    grad = np.zeros(2)
    info = np.zeros(2, dtype=np.int32)
    rc = lib().oracle_importance_discrete(_d(obs), _d(params), ctypes.c_double(float(rvr)),
                                          _d(rvp), ctypes.c_int(n), ctypes.c_int(nobs), _d(filt),
                                          _d(ll), _d(traj), _d(grad), _i(info))
    if rc != 0:
        raise MemoryError("oracle_importance_discrete failed")
    return dict(filt=filt, log_like=float(ll[0]), traj=traj, gradient=grad, traj_idx=int(info[0]),
                traj_oob=int(info[1]))


def stratified(rnd_sorted, n_data):
    rnd_sorted = _c64(rnd_sorted)
    m = rnd_sorted.shape[0]
    idx = np.empty(m, dtype=np.int32)
    rc = lib().oracle_stratified(_d(rnd_sorted), ctypes.c_int(m), ctypes.c_int(int(n_data)), _i(idx))
    if rc != 0:
        raise MemoryError("oracle_stratified failed")
    return idx


# --------------------------------------------------------------------------
# Python-level pieces of the path (NumPy restatements)
# --------------------------------------------------------------------------

def split_rvs_particle(rvs, nobs):
    """state/particle_methods/cython.py:89-91 -- note the FLAT split: the first
    NOBS flat entries feed the resampler, the remainder is rvp."""
    flat = np.asarray(rvs, dtype=np.float64).flatten()
    rv_r = norm.cdf(flat[0:nobs]).flatten()
    rv_p = flat[nobs:]
    return rv_r, rv_p


def split_rvs_importance(rvs):
    """state/importance_sampling/cython.py:82-83."""
    rvs = np.asarray(rvs, dtype=np.float64)
    rv_r = norm.cdf(rvs[:, 0][0])
    rv_p = rvs[:, 1:].flatten()
    return rv_r, rv_p


def smoother_post(gradient, hess1, hess2, compute_hessian):
    """state/particle_methods/cython.py:100-126: time-sum of the gradient with
    inf/nan zeroed, and the (Q9, scalar np.inner) Hessian assembly.
    Returns (grad_est[4], log_joint_hessian_estimate[4,4] or None)."""
    grad = np.array(gradient, dtype=np.float64).reshape((4, -1))
    grad[np.isinf(grad)] = 0.0
    grad[np.isnan(grad)] = 0.0
    grad_est = np.nansum(grad, axis=1)
    hess = None
    if compute_hessian:
        part1 = np.inner(grad_est, grad_est)
        part2 = np.array(hess1, dtype=np.float64).reshape((4, 4))
        part2 = part2 + np.array(hess2, dtype=np.float64).reshape((4, 4))
        hess = -(part1 - part2)
    return grad_est, hess


def subsample_indices(u, n_data):
    """state/direct/standard.py:75-76: sort(Phi(u)) -> stratified."""
    r = np.sort(norm.cdf(np.asarray(u, dtype=np.float64).flatten()))
    return stratified(r, n_data)


def logistic_loglike_gradient(beta, x_all, y_all, idx=None, compute_gradient=True,
                              compute_hessian=False):
    """models/logistic_regression.py:108-176 on the gathered rows.
    log-lik is NOT rescaled by n/m (reference behaviour)."""
    beta = np.asarray(beta, dtype=np.float64)
    if idx is None:
        x, y = x_all, y_all
    else:
        x, y = x_all[idx, :], y_all[idx]
    with np.errstate(over="ignore", divide="ignore", invalid="ignore"):
        xb = np.sum(beta * x, axis=1)
        eta = 1.0 / (1.0 + np.exp(-1.0 * xb))
        eta_1 = np.log(eta)
        eta_0 = np.log(1.0 - eta)
        eta_1[np.isinf(eta_1)] = 0.0
        eta_0[np.isinf(eta_0)] = 0.0
        log_like = np.sum(y * eta_1 + (1.0 - y) * eta_0)
        out = {"log_like": float(log_like)}
        if compute_gradient:
            grad_1 = x.T / (1.0 + np.exp(xb))
            grad_0 = -x.T / (1.0 + np.exp(-1.0 * xb))
            out["gradient"] = np.sum(y * grad_1 + (1.0 - y) * grad_0, axis=1)
        if compute_hessian:
            scale_0 = -(1.0 + np.exp(-xb)) ** (-2)
            scale_0 *= np.exp(-xb)
            scale_1 = -(1.0 + np.exp(xb)) ** (-2)
            scale_1 *= np.exp(xb)
            s = y * scale_1 + (1.0 - y) * scale_0
            out["hessian"] = -np.einsum("i,ij,ik->jk", s, x, x)
    return out


def crank_nicolson(u, xi, sigma_u):
    """parameter/mcmc/base_class.py:231-233."""
    mean = np.sqrt(1.0 - sigma_u ** 2) * u
    return mean + sigma_u * xi


# --------------------------------------------------------------------------
# synthetic data of BASELINE.json's shapes
# --------------------------------------------------------------------------

def simulate_sv(n_obs_plus_1, params=(0.2, 0.9, 0.4, -0.5), seed=87655678):
    """Returns from the SV-with-leverage model of models/stochastic_volatility.py:32-37:
    x_{t+1} = mu + phi (x_t - mu) + sigma_v v_t,  y_t = exp(x_t / 2) e_t,
    corr(v_t, e_t) = rho.  Returns y of length n_obs_plus_1."""
    mu, phi, sigmav, rho = params
    rs = np.random.RandomState(seed)
    n = int(n_obs_plus_1)
    x = np.zeros(n + 1)
    y = np.zeros(n)
    x[0] = mu + sigmav / np.sqrt(1.0 - phi * phi) * rs.normal()
    for t in range(n):
        e = rs.normal()
        v = rho * e + np.sqrt(1.0 - rho * rho) * rs.normal()
        y[t] = np.exp(0.5 * x[t]) * e
        x[t + 1] = mu + phi * (x[t] - mu) + sigmav * v
    return y


def simulate_re(n_obs=100, mu=1.0, sigma=0.2, seed=87655678):
    """Random-effects data as scripts/helper_random_effects.py:36-40 shapes it:
    x_i ~ N(mu, sigma^2), y_i ~ N(x_i, 1)."""
    rs = np.random.RandomState(seed)
    x = mu + sigma * rs.normal(size=n_obs)
    return x + rs.normal(size=n_obs)
