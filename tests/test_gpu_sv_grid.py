"""The grid kernel (csrc/sv_grid.cu) on its own: pmmh_sv_set_algorithm(6) runs it WITHOUT the
general-kernel fallback, so a pass here is a pass of that kernel (diag[PMMH_DIAG_KERNEL] == 5).

Parity against the CPU oracle (oracle_flps_sv_corr restating
state/particle_methods/stochastic_volatility.pyx:205-655): ancestors bit-exact, sorted generations,
log-likelihood rel <= 1e-10, filter / smoother means <= 1e-10, gradient <= 1e-9 * ||g||_inf
(the tolerances SURVEY.md 8(d) states); other lags; sizes that do not divide; N = 2^20 against the
general kernel (an independent device implementation); run-to-run bit reproducibility.
"""
import numpy as np
import pytest

import golden_inputs as gi
from helpers import first_mismatch_step, relerr, to_time_major

pytestmark = pytest.mark.gpu

DIAG_NEAR_TIES, DIAG_STATUS, DIAG_KERNEL, DIAG_INFO = 0, 2, 6, 7
GRID = 5   # diag[PMMH_DIAG_KERNEL] of the grid kernel


@pytest.fixture()
def grid_only():
    from pmmh_qn_b200 import kernels as K
    K.set_sv_algorithm(6)
    yield K
    K.set_sv_algorithm(0)


def _run(K, dev, obs, params, rvr, u_tm, lag, hist, ctas=0, hess=False):
    import torch
    out = K.flps_sv_corr(torch.from_numpy(obs).to(dev), torch.from_numpy(params).to(dev),
                         torch.from_numpy(rvr).to(dev), torch.from_numpy(u_tm).to(dev), lag=lag,
                         compute_hessian=hess, store_history=hist, ctas_per_problem=ctas)
    torch.cuda.synchronize()
    return {k: v.cpu().numpy() for k, v in out.items() if not k.startswith("_")}


def _inputs(n, nobs, seed):
    obs, params, rvr, rvp = gi.sv_inputs(n, nobs, seed)
    return obs, params, rvr[:nobs].copy(), rvp, to_time_major(rvp, n, nobs)


def _check_outputs(res, ref):
    assert abs(res["log_like"][0] - ref["log_like"]) <= 1e-10 * abs(ref["log_like"])
    assert relerr(res["filt"][0], ref["filt"]) <= 1e-10
    assert relerr(res["smo"][0], ref["smo"]) <= 1e-10
    assert relerr(res["traj"][0], ref["traj"]) <= 1e-12
    assert np.max(np.abs(res["gradient"][0] - ref["gradient"])) <= 1e-9 * np.max(np.abs(ref["gradient"]))


@pytest.mark.parametrize("n,nobs,ctas,seed", [(4096, 300, 4, 3), (20000, 120, 16, 3), (65536, 60, 64, 3),
                                              (6000, 200, 1, 4), (50000, 40, 148, 5), (5003, 80, 3, 1),
                                              (1000, 150, 7, 2)])
def test_grid_vs_oracle(cuda_dev, grid_only, n, nobs, ctas, seed):
    import oracle
    lag = 10
    obs, params, rvr, rvp, u = _inputs(n, nobs, seed)
    ref = oracle.flps_sv_corr(obs, params, rvr, rvp, n, lag, 0, dumps=True)
    res = _run(grid_only, cuda_dev, obs, params, rvr, u, lag, True, ctas)
    assert int(res["diag"][0, DIAG_KERNEL]) == GRID, res["diag"]
    assert int(res["diag"][0, DIAG_STATUS]) == 0, "abandoned: info %x" % int(res["diag"][0, DIAG_INFO])
    step = first_mismatch_step(res["A"][0][1:], ref["A"][1:])
    assert step is None, "ancestors differ first at time %d (near ties reported: %d)" % (
        step + 1, int(res["diag"][0, DIAG_NEAR_TIES]))
    assert relerr(res["X"][0], ref["X"]) <= 1e-12
    _check_outputs(res, ref)


@pytest.mark.parametrize("lag", [2, 3, 4, 6, 7, 9])
def test_grid_other_lags(cuda_dev, grid_only, lag):
    import oracle
    n, nobs = 5000, 80
    obs, params, rvr, rvp, u = _inputs(n, nobs, 1)
    ref = oracle.flps_sv_corr(obs, params, rvr, rvp, n, lag, 0, dumps=True)
    res = _run(grid_only, cuda_dev, obs, params, rvr, u, lag, False, 5)
    assert int(res["diag"][0, DIAG_KERNEL]) == GRID and int(res["diag"][0, DIAG_STATUS]) == 0
    _check_outputs(res, ref)


@pytest.mark.parametrize("n,nobs,ctas,lag,seed", [(4096, 300, 4, 10, 3), (20000, 120, 16, 10, 3), (5003, 80, 3, 7, 1),
                                                  (1000, 150, 7, 10, 2), (5000, 60, 5, 2, 1), (5000, 60, 5, 3, 1),
                                                  (65536, 45, 148, 10, 5), (300, 400, 2, 10, 4)])
def test_grid_hessian_vs_oracle(cuda_dev, grid_only, n, nobs, ctas, lag, seed):
    """Hessian branch (stochastic_volatility.pyx:361-390,472-534,564-626 with Q7 / Q8) on the grid kernel's
    second instantiation: hessian1 / hessian2 <= 1e-8 * max|H| (the tolerance of the other Hessian tests),
    everything else as without the Hessian.  n < n_obs makes Q7 read real (sorted / unsorted) values."""
    import oracle
    obs, params, rvr, rvp, u = _inputs(n, nobs, seed)
    ref = oracle.flps_sv_corr(obs, params, rvr, rvp, n, lag, 1, dumps=True)
    res = _run(grid_only, cuda_dev, obs, params, rvr, u, lag, True, ctas, hess=True)
    assert int(res["diag"][0, DIAG_KERNEL]) == GRID, res["diag"]
    assert int(res["diag"][0, DIAG_STATUS]) == 0, "abandoned: info %x" % int(res["diag"][0, DIAG_INFO])
    assert first_mismatch_step(res["A"][0][1:], ref["A"][1:]) is None
    _check_outputs(res, ref)
    for k in ("hess1", "hess2"):
        h = res[k][0].reshape(4, 4)
        assert np.max(np.abs(h - ref[k].reshape(4, 4))) <= 1e-8 * np.max(np.abs(ref[k])), (k, h, ref[k])


@pytest.mark.parametrize("n", [1 << 20, 1000003])
def test_grid_vs_general_full_size(cuda_dev, n):
    """BASELINE size N = 2^20 (T cut to 40 steps) and a size that is not a power of two: the grid
    kernel and the general kernel agree on every output."""
    import torch
    from pmmh_qn_b200 import kernels as K
    nobs, lag = 41, 10
    dev = cuda_dev
    g = torch.Generator(device=dev)
    g.manual_seed(7)
    obs = torch.from_numpy(gi.sv_obs(nobs)).to(dev)
    params = torch.tensor([[0.2, 0.9, 0.4, -0.5]], dtype=torch.float64, device=dev)
    u = torch.randn((1, nobs, n), dtype=torch.float64, device=dev, generator=g)
    rvr = torch.rand((1, nobs), dtype=torch.float64, device=dev, generator=g)
    outs = {}
    try:
        for algo in (1, 6):
            K.set_sv_algorithm(algo)
            o = K.flps_sv_corr(obs, params, rvr, u, lag=lag, compute_hessian=False)
            torch.cuda.synchronize()
            outs[algo] = {k: v.cpu().numpy() for k, v in o.items() if not k.startswith("_")}
    finally:
        K.set_sv_algorithm(0)
    a, b = outs[1], outs[6]
    assert int(a["diag"][0, DIAG_KERNEL]) == 1 and int(b["diag"][0, DIAG_KERNEL]) == GRID
    assert int(b["diag"][0, DIAG_STATUS]) == 0, "abandoned: info %x" % int(b["diag"][0, DIAG_INFO])
    assert abs(a["log_like"][0] - b["log_like"][0]) <= 1e-12 * abs(a["log_like"][0])
    assert relerr(b["filt"][0], a["filt"][0]) <= 1e-11
    assert relerr(b["smo"][0], a["smo"][0]) <= 1e-10
    assert relerr(b["traj"][0], a["traj"][0]) <= 1e-12
    assert np.max(np.abs(a["gradient"][0] - b["gradient"][0])) <= 1e-9 * np.max(np.abs(a["gradient"][0]))


def test_grid_is_bit_reproducible(cuda_dev, grid_only):
    n, nobs = 200000, 50
    obs, params, rvr, rvp, u = _inputs(n, nobs, 5)
    r1 = _run(grid_only, cuda_dev, obs, params, rvr, u, 10, False)
    r2 = _run(grid_only, cuda_dev, obs, params, rvr, u, 10, False)
    assert int(r1["diag"][0, DIAG_KERNEL]) == GRID and int(r1["diag"][0, DIAG_STATUS]) == 0
    for k in ("log_like", "filt", "smo", "gradient", "traj"):
        assert np.array_equal(r1[k], r2[k]), k


def test_grid_history_invariants_large(cuda_dev, grid_only):
    """N = 2^19: every stored generation is sorted and the composed ancestors are valid sorted
    positions whose multiset is the resampling outcome (monotone before the sort)."""
    n, nobs = 1 << 19, 24
    obs, params, rvr, rvp, u = _inputs(n, nobs, 2)
    res = _run(grid_only, cuda_dev, obs, params, rvr, u, 10, True)
    assert int(res["diag"][0, DIAG_KERNEL]) == GRID and int(res["diag"][0, DIAG_STATUS]) == 0
    X, A = res["X"][0], res["A"][0]
    assert np.all(np.diff(X, axis=1) >= 0.0)
    assert A.min() >= 0 and A.max() < n
    for t in range(1, nobs):
        counts = np.bincount(A[t], minlength=n)
        assert counts.sum() == n


def test_grid_automatic_selection_and_fallback(cuda_dev):
    """algorithm 0 picks the grid kernel for one large problem; a degenerate cloud (all auxiliary
    variables equal: every child of a parent has the same value) is abandoned and re-run by the
    general kernel."""
    import torch
    from pmmh_qn_b200 import kernels as K
    K.set_sv_algorithm(0)
    n, nobs = 1 << 17, 30
    obs, params, rvr, rvp, u = _inputs(n, nobs, 0)
    res = _run(K, cuda_dev, obs, params, rvr, u, 10, False)
    assert int(res["diag"][0, DIAG_KERNEL]) == GRID and int(res["diag"][0, DIAG_STATUS]) == 0
    u0 = np.zeros_like(u)
    res0 = _run(K, cuda_dev, obs, params, rvr, u0, 10, False)
    assert int(res0["diag"][0, DIAG_KERNEL]) == 1, res0["diag"]


@pytest.mark.gpu
@pytest.mark.parametrize("pinned,n,nobs", [(True, 65536, 300), (False, 100000, 70), (True, 1 << 20, 30)])
def test_grid_host_streamed_rvs_equals_resident(cuda_dev, pinned, n, nobs):
    """pmmh_flps_sv_corr_streamed on the grid kernel: the copy engine lays the reference's (NOBS, N+1)
    host array down in particle-major chunks of 256 time steps while the kernel runs and polls one flag
    per step; bit-identical to upload -> time-major array -> pmmh_flps_sv_corr."""
    import torch
    from helpers import to_time_major
    from pmmh_qn_b200 import kernels as K
    lag, dev = 10, cuda_dev
    rs = np.random.RandomState(33)
    if pinned:
        host = torch.empty((nobs, n + 1), dtype=torch.float64, pin_memory=True)
        rvs = host.numpy()
        rvs[...] = rs.normal(size=(nobs, n + 1))
    else:
        rvs = rs.normal(size=(nobs, n + 1))
    assert K.sv_streamed_eligible(nobs, n, lag)
    obs = torch.from_numpy(gi.sv_obs(nobs)).to(dev)
    params = torch.tensor([[0.2, 0.9, 0.4, -0.5]], dtype=torch.float64, device=dev)
    rvr_h, rvp = gi.split_particle(rvs, nobs)
    rvr = torch.from_numpy(rvr_h).to(dev)
    ws, st = K.Workspace(), K.Workspace()
    a = K.flps_sv_corr_streamed(rvs, obs, params, rvr, nobs, n, lag=lag, workspace=ws, stage=st)
    a2 = K.flps_sv_corr_streamed(rvs, obs, params, rvr, nobs, n, lag=lag, workspace=ws, stage=st)   # back to back
    torch.cuda.synchronize()
    assert int(a["diag"][0, DIAG_KERNEL]) == GRID and int(a["diag"][0, DIAG_STATUS]) == 0
    u = torch.from_numpy(to_time_major(rvp, n, nobs)).to(dev)
    b = K.flps_sv_corr(obs, params, rvr, u, lag=lag, compute_hessian=False)
    torch.cuda.synchronize()
    assert int(b["diag"][0, DIAG_KERNEL]) == GRID
    for k in ("log_like", "filt", "smo", "gradient", "traj"):
        assert torch.equal(a[k], b[k]), k
        assert torch.equal(a2[k], b[k]), k


@pytest.mark.parametrize("logn,nobs,seed", [(20, 25, 7), (22, 21, 8)])
def test_large_n_against_the_oracle(cuda_dev, logn, nobs, seed):
    """The sizes of BASELINE configs[1] (N = 2^20, the headline) and configs[4] (N = 2^22) themselves against
    the CPU oracle over a short series: automatic kernel selection (grid kernel at 2^20, streaming kernels
    with path storage at 2^22).  Ancestors are compared exactly; if a generation differs, the first one must
    come with a counted tie (at this N the reference's sequential cumulative sums and any
    parallel sum round differently about once per time step)."""
    import oracle
    import torch
    from pmmh_qn_b200 import kernels as K
    n, lag, dev = 1 << logn, 10, cuda_dev
    g = torch.Generator(device=dev)
    g.manual_seed(seed)
    u = torch.randn((nobs, n), dtype=torch.float64, device=dev, generator=g)
    rvr = torch.rand((nobs,), dtype=torch.float64, device=dev, generator=g)
    obs = gi.sv_obs(nobs)
    params = np.array(gi.SV_PARAM_SETS[0])
    store = logn <= 20
    out = K.flps_sv_corr(torch.from_numpy(obs).to(dev), torch.from_numpy(params).to(dev), rvr, u, lag=lag,
                         compute_hessian=False, store_history=store)
    torch.cuda.synchronize()
    d = out["diag"][0].cpu().numpy()
    assert int(d[DIAG_STATUS]) == 0 and int(d[DIAG_KERNEL]) == (GRID if logn <= 20 else 4)
    rvp = np.ascontiguousarray(u.cpu().numpy().T).reshape(-1)
    ref = oracle.flps_sv_corr(obs, params, rvr.cpu().numpy(), rvp, n, lag, 0, dumps=store)
    ties = int(d[4]) if logn <= 20 else int(d[DIAG_NEAR_TIES])
    exact = True
    if store:
        step = first_mismatch_step(out["A"][0].cpu().numpy()[1:], ref["A"][1:])
        exact = step is None
        assert exact or ties > 0, "ancestors differ at time %d without a counted tie" % (step + 1)
    ll = float(out["log_like"][0])
    assert abs(ll - ref["log_like"]) <= 1e-10 * abs(ref["log_like"])
    assert relerr(out["filt"][0].cpu().numpy(), ref["filt"]) <= 1e-9
    gtol = 1e-9 if exact and store else 1e-4          # a flipped neighbour moves the fixed-lag means by O(1/N)
    gr = out["gradient"][0].cpu().numpy()
    assert np.max(np.abs(gr - ref["gradient"])) <= gtol * np.max(np.abs(ref["gradient"]))
    assert relerr(out["smo"][0].cpu().numpy(), ref["smo"]) <= gtol
