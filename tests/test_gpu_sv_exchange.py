"""The exchange kernel (csrc/sv_fast.cu) on its own: pmmh_sv_set_algorithm(2) runs it WITHOUT the
general-kernel fallback, so a pass here is a pass of that kernel (diag[PMMH_DIAG_KERNEL] == 2).

Parity against the CPU oracle (ancestors bit-exact, near-ties counted), against the general
kernel at N = 2^20 (two independent device implementations), run-to-run bit reproducibility,
other lags, batches over several teams, and size-independent invariants of the stored history.
"""
import numpy as np
import pytest

import golden_inputs as gi
from helpers import first_mismatch_step, relerr, to_time_major

pytestmark = pytest.mark.gpu

DIAG_NEAR_TIES, DIAG_STATUS, DIAG_KERNEL = 0, 2, 6


@pytest.fixture()
def exchange_only():
    from pmmh_qn_b200 import kernels as K
    K.set_sv_algorithm(2)
    yield K
    K.set_sv_algorithm(0)


def _run(K, dev, obs, params, rvr, u_tm, lag, hist, ctas=0):
    import torch
    out = K.flps_sv_corr(torch.from_numpy(obs).to(dev), torch.from_numpy(params).to(dev),
                         torch.from_numpy(rvr).to(dev), torch.from_numpy(u_tm).to(dev), lag=lag,
                         compute_hessian=False, store_history=hist, ctas_per_problem=ctas)
    torch.cuda.synchronize()
    return {k: v.cpu().numpy() for k, v in out.items() if not k.startswith("_")}


def _inputs(n, nobs, seed):
    obs, params, rvr, rvp = gi.sv_inputs(n, nobs, seed)
    return obs, params, rvr[:nobs].copy(), rvp, to_time_major(rvp, n, nobs)


@pytest.mark.parametrize("n,nobs,ctas", [(4096, 300, 4), (20000, 120, 16), (65536, 60, 64),
                                         (6000, 200, 1), (50000, 40, 148)])
def test_exchange_vs_oracle(cuda_dev, exchange_only, n, nobs, ctas):
    import oracle
    lag = 10
    obs, params, rvr, rvp, u = _inputs(n, nobs, 3)
    ref = oracle.flps_sv_corr(obs, params, rvr, rvp, n, lag, 0, dumps=True)
    res = _run(exchange_only, cuda_dev, obs, params, rvr, u, lag, True, ctas)
    assert int(res["diag"][0, DIAG_KERNEL]) == 2 and int(res["diag"][0, DIAG_STATUS]) == 0
    step = first_mismatch_step(res["A"][0][1:], ref["A"][1:])
    assert step is None, "ancestors differ first at time %d (near ties reported: %d)" % (
        step + 1, int(res["diag"][0, DIAG_NEAR_TIES]))
    assert relerr(res["X"][0], ref["X"]) <= 1e-12
    assert abs(res["log_like"][0] - ref["log_like"]) <= 1e-10 * abs(ref["log_like"])
    assert relerr(res["filt"][0], ref["filt"]) <= 1e-10
    assert relerr(res["smo"][0], ref["smo"]) <= 1e-10
    assert relerr(res["traj"][0], ref["traj"]) <= 1e-12
    assert np.max(np.abs(res["gradient"][0] - ref["gradient"])) <= 1e-9 * np.max(np.abs(ref["gradient"]))


@pytest.mark.parametrize("lag", [2, 3, 4, 6, 7, 11, 13])
def test_exchange_other_lags(cuda_dev, exchange_only, lag):
    import oracle
    n, nobs = 5000, 80
    obs, params, rvr, rvp, u = _inputs(n, nobs, 1)
    ref = oracle.flps_sv_corr(obs, params, rvr, rvp, n, lag, 0, dumps=True)
    res = _run(exchange_only, cuda_dev, obs, params, rvr, u, lag, False, 5)
    assert int(res["diag"][0, DIAG_KERNEL]) == 2 and int(res["diag"][0, DIAG_STATUS]) == 0
    assert abs(res["log_like"][0] - ref["log_like"]) <= 1e-10 * abs(ref["log_like"])
    assert relerr(res["smo"][0], ref["smo"]) <= 1e-10
    assert np.max(np.abs(res["gradient"][0] - ref["gradient"])) <= 1e-9 * np.max(np.abs(ref["gradient"]))


def test_exchange_vs_general_full_size(cuda_dev):
    """BASELINE size N = 2^20 (T cut to 40 steps): the exchange kernel and the general kernel
    agree on every output; ancestors are compared through the outputs they determine."""
    import torch
    from pmmh_qn_b200 import kernels as K
    n, nobs, lag = 1 << 20, 41, 10
    dev = cuda_dev
    g = torch.Generator(device=dev)
    g.manual_seed(7)
    obs = torch.from_numpy(gi.sv_obs(nobs)).to(dev)
    params = torch.tensor([[0.2, 0.9, 0.4, -0.5]], dtype=torch.float64, device=dev)
    u = torch.randn((1, nobs, n), dtype=torch.float64, device=dev, generator=g)
    rvr = torch.rand((1, nobs), dtype=torch.float64, device=dev, generator=g)
    outs = {}
    try:
        for algo in (1, 2):
            K.set_sv_algorithm(algo)
            o = K.flps_sv_corr(obs, params, rvr, u, lag=lag, compute_hessian=False)
            torch.cuda.synchronize()
            outs[algo] = {k: v.cpu().numpy() for k, v in o.items() if not k.startswith("_")}
    finally:
        K.set_sv_algorithm(0)
    a, b = outs[1], outs[2]
    assert int(a["diag"][0, DIAG_KERNEL]) == 1 and int(b["diag"][0, DIAG_KERNEL]) == 2
    assert int(b["diag"][0, DIAG_STATUS]) == 0
    assert abs(a["log_like"][0] - b["log_like"][0]) <= 1e-12 * abs(a["log_like"][0])
    assert relerr(b["filt"][0], a["filt"][0]) <= 1e-11
    assert relerr(b["smo"][0], a["smo"][0]) <= 1e-10
    assert relerr(b["traj"][0], a["traj"][0]) <= 1e-12
    assert np.max(np.abs(a["gradient"][0] - b["gradient"][0])) <= 1e-9 * np.max(np.abs(a["gradient"][0]))


def test_exchange_is_bit_reproducible(cuda_dev, exchange_only):
    n, nobs = 200000, 50
    obs, params, rvr, rvp, u = _inputs(n, nobs, 5)
    r1 = _run(exchange_only, cuda_dev, obs, params, rvr, u, 10, False)
    r2 = _run(exchange_only, cuda_dev, obs, params, rvr, u, 10, False)
    assert int(r1["diag"][0, DIAG_KERNEL]) == 2 and int(r1["diag"][0, DIAG_STATUS]) == 0
    for k in ("log_like", "filt", "smo", "gradient", "traj"):
        assert np.array_equal(r1[k], r2[k]), k


def test_exchange_history_invariants_large(cuda_dev, exchange_only):
    """N = 2^19: every stored generation is sorted, ancestors are valid positions, and the
    composed one-step ancestors of the sorted children follow the parents' order only through
    the sort (no value may be lost: the multiset of children per parent sums to N)."""
    n, nobs = 1 << 19, 14
    obs, params, rvr, rvp, u = _inputs(n, nobs, 2)
    res = _run(exchange_only, cuda_dev, obs, params, rvr, u, 10, True)
    assert int(res["diag"][0, DIAG_KERNEL]) == 2 and int(res["diag"][0, DIAG_STATUS]) == 0
    X, A = res["X"][0], res["A"][0]
    assert np.all(np.diff(X, axis=1) >= 0.0), "a generation is not sorted"
    assert A.min() >= 0 and A.max() < n
    for t in range(1, nobs):
        counts = np.bincount(A[t], minlength=n)
        assert counts.sum() == n
        # systematic resampling: offspring numbers differ from N w by less than one
        assert counts.max() <= n


def test_exchange_batch_over_teams(cuda_dev, exchange_only):
    """B problems spread over several teams in one launch == B single launches, bit for bit."""
    import torch
    K = exchange_only
    n, nobs, lag, B = 3000, 90, 10, 7
    obs = gi.sv_obs(nobs)
    rs = np.random.RandomState(11)
    params = np.array(gi.SV_PARAM_SETS[0]) + 0.02 * rs.normal(size=(B, 4))
    rvs = rs.normal(size=(B, nobs, n + 1))
    rvr = np.zeros((B, nobs))
    u = np.zeros((B, nobs, n))
    for b in range(B):
        r, p = gi.split_particle(rvs[b], nobs)
        rvr[b] = r
        u[b] = to_time_major(p, n, nobs)
    dev = cuda_dev
    t = lambda x: torch.from_numpy(x).to(dev)   # noqa: E731
    out = K.flps_sv_corr(t(obs), t(params), t(rvr), t(u), lag=lag, ctas_per_problem=3)
    torch.cuda.synchronize()
    assert np.all(out["diag"][:, DIAG_KERNEL].cpu().numpy() == 2)
    assert np.all(out["diag"][:, DIAG_STATUS].cpu().numpy() == 0)
    for b in range(B):
        single = K.flps_sv_corr(t(obs), t(params[b:b + 1]), t(rvr[b:b + 1]), t(u[b:b + 1]), lag=lag,
                                ctas_per_problem=3)
        torch.cuda.synchronize()
        for k in ("log_like", "filt", "smo", "gradient", "traj"):
            assert torch.equal(out[k][b], single[k][0]), (b, k)


@pytest.mark.parametrize("pinned", [True, False])
def test_streamed_host_rvs_equals_resident(cuda_dev, pinned):
    """pmmh_flps_sv_corr_streamed (copies overlap the kernel, chunked particle-major staging) gives
    bit-identical outputs to the upload -> split_rvs -> pmmh_flps_sv_corr path."""
    import torch
    from pmmh_qn_b200 import kernels as K
    n, nobs, lag = 65536, 150, 10
    dev = cuda_dev
    rs = np.random.RandomState(21)
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
    K.set_sv_algorithm(2)      # (automatic selection would take the grid kernel at this size)
    try:
        a = K.flps_sv_corr_streamed(rvs, obs, params, rvr, nobs, n, lag=lag)
        torch.cuda.synchronize()
        assert int(a["diag"][0, DIAG_KERNEL]) == 2 and int(a["diag"][0, DIAG_STATUS]) == 0
        u = torch.from_numpy(to_time_major(rvp, n, nobs)).to(dev)
        b = K.flps_sv_corr(obs, params, rvr, u, lag=lag, compute_hessian=False)
        torch.cuda.synchronize()
    finally:
        K.set_sv_algorithm(0)
    for k in ("log_like", "filt", "smo", "gradient", "traj"):
        assert torch.equal(a[k], b[k]), k
    # twice in a row on the same staging buffer (the second call must wait for the first)
    ws, st = K.Workspace(), K.Workspace()
    K.set_sv_algorithm(2)
    try:
        c1 = K.flps_sv_corr_streamed(rvs, obs, params, rvr, nobs, n, lag=lag, workspace=ws, stage=st)
        c2 = K.flps_sv_corr_streamed(rvs, obs, params, rvr, nobs, n, lag=lag, workspace=ws, stage=st)
        torch.cuda.synchronize()
    finally:
        K.set_sv_algorithm(0)
    for k in ("log_like", "gradient"):
        assert torch.equal(c1[k], a[k]) and torch.equal(c2[k], a[k]), k


# ------------------------------------------------------------------ chain kernel (one CTA per problem)
@pytest.fixture()
def chain_only():
    from pmmh_qn_b200 import kernels as K
    K.set_sv_algorithm(3)
    yield K
    K.set_sv_algorithm(0)


@pytest.mark.parametrize("n,nobs,lag", [(75, 361, 10), (1024, 301, 10), (3000, 200, 10), (4096, 1001, 10),
                                        (4096, 120, 2), (2500, 120, 3), (2500, 120, 6), (2500, 120, 7)])
def test_chain_vs_oracle(cuda_dev, chain_only, n, nobs, lag):
    import oracle
    obs, params, rvr, rvp, u = _inputs(n, nobs, 4)
    ref = oracle.flps_sv_corr(obs, params, rvr, rvp, n, lag, 0, dumps=True)
    res = _run(chain_only, cuda_dev, obs, params, rvr, u, lag, True)
    assert int(res["diag"][0, DIAG_KERNEL]) == 3 and int(res["diag"][0, DIAG_STATUS]) == 0
    step = first_mismatch_step(res["A"][0][1:], ref["A"][1:])
    assert step is None, "ancestors differ first at time %d" % (step + 1)
    assert relerr(res["X"][0], ref["X"]) <= 1e-12
    assert abs(res["log_like"][0] - ref["log_like"]) <= 1e-10 * abs(ref["log_like"])
    assert relerr(res["filt"][0], ref["filt"]) <= 1e-10
    assert relerr(res["smo"][0], ref["smo"]) <= 1e-10
    assert relerr(res["traj"][0], ref["traj"]) <= 1e-12
    assert np.max(np.abs(res["gradient"][0] - ref["gradient"])) <= 1e-9 * np.max(np.abs(ref["gradient"]))


@pytest.mark.parametrize("n,nobs,lag,seed", [(75, 361, 10, 0), (200, 120, 10, 1), (37, 50, 10, 1), (64, 40, 4, 0),
                                             (1024, 1001, 10, 0), (4096, 1001, 10, 0), (3000, 90, 10, 2),
                                             (4096, 60, 10, 1), (2500, 120, 3, 2), (2500, 120, 2, 1)])
def test_chain_hessian_branch_vs_oracle(cuda_dev, chain_only, n, nobs, lag, seed):
    """The Hessian branch (:361-390 alpha recursion with Q7 / Q8, :472-534, :564-626) on the chain kernel,
    no fallback pass: hess1 / hess2 within 1e-8 of the oracle, everything else as without it."""
    import oracle
    import torch
    obs, params, rvr, rvp = gi.sv_inputs(n, nobs, seed)
    ref = oracle.flps_sv_corr(obs, params, rvr, rvp, n, lag, 1, dumps=True)
    dev = cuda_dev
    out = chain_only.flps_sv_corr(torch.from_numpy(obs).to(dev), torch.from_numpy(params).to(dev),
                                  torch.from_numpy(rvr[:nobs].copy()).to(dev),
                                  torch.from_numpy(to_time_major(rvp, n, nobs)).to(dev), lag=lag,
                                  compute_hessian=True, store_history=True)
    torch.cuda.synchronize()
    res = {k: v.cpu().numpy() for k, v in out.items() if not k.startswith("_")}
    assert int(res["diag"][0, DIAG_KERNEL]) == 3 and int(res["diag"][0, DIAG_STATUS]) == 0
    assert first_mismatch_step(res["A"][0][1:], ref["A"][1:]) is None
    assert abs(res["log_like"][0] - ref["log_like"]) <= 1e-10 * abs(ref["log_like"])
    assert np.max(np.abs(res["gradient"][0] - ref["gradient"])) <= 1e-9 * np.max(np.abs(ref["gradient"]))
    for k in ("hess1", "hess2"):
        assert np.max(np.abs(res[k][0] - ref[k])) <= 1e-8 * np.max(np.abs(ref[k])), k
        assert np.array_equal(res[k][0], res[k][0].T)


def test_chain_hessian_batch_and_golden(cuda_dev, chain_only, golden):
    """Batch of chains with the Hessian branch: every problem equals its single launch bit for bit; the
    shipped size against the compiled reference's own outputs (tests/golden)."""
    import torch
    K, dev = chain_only, cuda_dev
    n, nobs, lag, B = 1500, 80, 10, 160
    obs = gi.sv_obs(nobs)
    rs = np.random.RandomState(31)
    params = np.array(gi.SV_PARAM_SETS[0]) + 0.02 * rs.normal(size=(B, 4))
    g = torch.Generator(device=dev)
    g.manual_seed(5)
    u = torch.randn((B, nobs, n), dtype=torch.float64, device=dev, generator=g)
    rvr = torch.rand((B, nobs), dtype=torch.float64, device=dev, generator=g)
    t = lambda x: torch.from_numpy(x).to(dev)   # noqa: E731
    out = K.flps_sv_corr(t(obs), t(params), rvr, u, lag=lag, compute_hessian=True)
    torch.cuda.synchronize()
    assert np.all(out["diag"][:, DIAG_KERNEL].cpu().numpy() == 3) and np.all(out["diag"][:, DIAG_STATUS].cpu().numpy() == 0)
    for b in (0, 77, 159):
        single = K.flps_sv_corr(t(obs), t(params[b:b + 1]), rvr[b:b + 1].contiguous(), u[b:b + 1].contiguous(),
                                lag=lag, compute_hessian=True)
        torch.cuda.synchronize()
        for k in ("log_like", "gradient", "hess1", "hess2"):
            assert torch.equal(out[k][b], single[k][0]), (b, k)
    gg = golden["sv_kernels"]
    for (n, nobs, lag, seeds) in gi.SV_KERNEL_CASES:
        for seed in seeds:
            obs, params, rvr, rvp = gi.sv_inputs(n, nobs, seed)
            o = K.flps_sv_corr(t(obs), t(params), t(rvr[:nobs].copy()), t(to_time_major(rvp, n, nobs)), lag=lag,
                               compute_hessian=True)
            torch.cuda.synchronize()
            assert int(o["diag"][0, DIAG_KERNEL]) == 3
            pre = "flps_n%d_t%d_l%d_s%d_h1_" % (n, nobs, lag, seed)
            for k in ("hess1", "hess2"):
                href = gg[pre + k].reshape(4, 4)
                assert np.max(np.abs(o[k][0].cpu().numpy() - href)) <= 1e-8 * np.max(np.abs(href)), pre + k


def test_chain_batch_equals_singles_and_general(cuda_dev, chain_only):
    """A batch of chains (more problems than CTAs) == single launches bit for bit, and agrees with
    the general kernel within the tolerances."""
    import torch
    K = chain_only
    n, nobs, lag, B = 2048, 100, 10, 200
    obs = gi.sv_obs(nobs)
    rs = np.random.RandomState(13)
    params = np.array(gi.SV_PARAM_SETS[0]) + 0.02 * rs.normal(size=(B, 4))
    dev = cuda_dev
    g = torch.Generator(device=dev)
    g.manual_seed(3)
    u = torch.randn((B, nobs, n), dtype=torch.float64, device=dev, generator=g)
    rvr = torch.rand((B, nobs), dtype=torch.float64, device=dev, generator=g)
    t = lambda x: torch.from_numpy(x).to(dev)   # noqa: E731
    out = K.flps_sv_corr(t(obs), t(params), rvr, u, lag=lag)
    torch.cuda.synchronize()
    assert np.all(out["diag"][:, DIAG_KERNEL].cpu().numpy() == 3)
    assert np.all(out["diag"][:, DIAG_STATUS].cpu().numpy() == 0)
    for b in (0, 57, 148, 199):
        single = K.flps_sv_corr(t(obs), t(params[b:b + 1]), rvr[b:b + 1].contiguous(), u[b:b + 1].contiguous(), lag=lag)
        torch.cuda.synchronize()
        for k in ("log_like", "filt", "smo", "gradient", "traj"):
            assert torch.equal(out[k][b], single[k][0]), (b, k)
    K.set_sv_algorithm(1)
    gen = K.flps_sv_corr(t(obs), t(params), rvr, u, lag=lag)
    torch.cuda.synchronize()
    a, b_ = out["log_like"].cpu().numpy(), gen["log_like"].cpu().numpy()
    assert np.max(np.abs(a - b_) / np.abs(b_)) <= 1e-12
    ga, gb = out["gradient"].cpu().numpy(), gen["gradient"].cpu().numpy()
    assert np.max(np.abs(ga - gb)) <= 1e-9 * np.max(np.abs(gb))
