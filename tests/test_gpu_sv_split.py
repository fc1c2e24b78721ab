"""GPU parity of the split particle filter / smoother (pmmh_svsplit_*, BASELINE configs[4]) against
the CPU oracle.  The ranks run in ONE process on one GPU (LocalComm): same device phases and
the same partition logic as the NCCL path, exchanges as device copies.

Bar: every sorted generation, gathered over the ranks, equals the oracle's to 1e-12 (that is
ancestors + propagation + sort: one wrong ancestor would move a value by ~1e-6); log-likelihood rel <= 1e-10, filter / smoother means rel <= 1e-10,
gradient <= 1e-9 * max|g| (summation order differs from the reference's sequential loops).
"""
import numpy as np
import pytest

import golden_inputs as gi
from helpers import relerr, to_time_major

pytestmark = pytest.mark.gpu


def _run(dev, world, obs, params, rvr, rvp, n, nobs, lag, cap=None, capc=None, history=True):
    import torch
    from pmmh_qn_b200.state.particle_methods import split as SP
    u = torch.from_numpy(to_time_major(rvp, n, nobs)).to(dev)
    rvr_d = torch.from_numpy(rvr[:nobs].copy()).to(dev)
    out = SP.run_split_smoother(SP.LocalComm(world), obs, params, n, lag, rvr_d, u_d=u, device=dev,
                                cap=cap, capc=capc, keep_history=history)
    torch.cuda.synchronize()
    return out


def _check(out, ref, nobs, lag, history=True):
    for o in out["per_rank"]:
        assert o["diag"][2] == 0, "status %d" % o["diag"][2]
    if history:
        for t in range(nobs):
            x = np.concatenate([c.cpu().numpy() for c in out["x_hist"][t]])
            assert x.shape[0] == ref["X"].shape[1]
            # a single wrong ancestor moves one value by ~ the parent spacing (>> 1e-12); libm and
            # CUDA exp() differ in the last ulp, so bit equality of the values is not the bar
            assert relerr(x, ref["X"][t]) <= 1e-12, "generation %d differs from the oracle" % t
    ll = float(out["log_like"].item())
    assert abs(ll - ref["log_like"]) <= 1e-10 * abs(ref["log_like"]), (ll, ref["log_like"])
    assert relerr(out["filt"].cpu().numpy(), ref["filt"]) <= 1e-10
    assert relerr(out["traj"].cpu().numpy(), ref["traj"]) <= 1e-12
    if lag:
        assert relerr(out["smo"].cpu().numpy(), ref["smo"]) <= 1e-10
        g, gr = out["gradient"].cpu().numpy(), ref["gradient"]
        assert np.max(np.abs(g - gr)) <= 1e-9 * np.max(np.abs(gr))
    # every rank assembles the same outputs
    for o in out["per_rank"][1:]:
        assert float(o["log_like"].item()) == ll
        assert np.array_equal(o["gradient"].cpu().numpy(), out["gradient"].cpu().numpy())


@pytest.mark.parametrize("world", [1, 2, 3, 4])
@pytest.mark.parametrize("n,nobs,lag,seed", [(2000, 61, 10, 0), (777, 45, 4, 1), (5000, 40, 10, 2)])
def test_split_smoother_vs_oracle(cuda_dev, world, n, nobs, lag, seed):
    import oracle
    obs, params, rvr, rvp = gi.sv_inputs(n, nobs, seed)
    ref = oracle.flps_sv_corr(obs, params, rvr, rvp, n, lag, 0, dumps=True)
    out = _run(cuda_dev, world, obs, params, rvr, rvp, n, nobs, lag, cap=n, capc=n)
    _check(out, ref, nobs, lag)


@pytest.mark.parametrize("world", [2, 8])
def test_split_filter_only_and_balance(cuda_dev, world):
    """lag = 0: records are the bare values.  Arrivals per rank stay within one histogram bin of
    N / world."""
    import oracle
    n, nobs = 1 << 17, 40
    obs, params, rvr, rvp = gi.sv_inputs(n, nobs, 3)
    ref = oracle.flps_sv_corr(obs, params, rvr, rvp, n, 10, 0, dumps=True)
    out = _run(cuda_dev, world, obs, params, rvr, rvp, n, nobs, 0)
    _check(out, ref, nobs, 0)
    counts = out["counts"][1:]
    assert counts.sum(axis=1).tolist() == [n] * (nobs - 1)
    assert counts.max() <= n / world * 1.05 + 64


def test_split_smoother_default_capacities_large(cuda_dev):
    """N = 2^19 over 4 ranks, T = 30, default capacities, gradient: against the oracle."""
    import oracle
    n, nobs, lag = 1 << 19, 31, 10
    obs, params, rvr, rvp = gi.sv_inputs(n, nobs, 5)
    ref = oracle.flps_sv_corr(obs, params, rvr, rvp, n, lag, 0, dumps=True)
    out = _run(cuda_dev, 4, obs, params, rvr, rvp, n, nobs, lag)
    _check(out, ref, nobs, lag)


def test_philox_stream_equals_the_array(cuda_dev):
    """u regenerated from the Philox stream inside the kernel == the same stream materialised."""
    import torch
    from pmmh_qn_b200 import kernels as K
    from pmmh_qn_b200.state.particle_methods import split as SP
    n, nobs, lag = 30000, 50, 10
    obs = gi.sv_obs(nobs)
    params = np.array(gi.SV_PARAM_SETS[0], dtype=np.float64)
    ph = SP.PhiloxRVS(seed=1234, offset=77)
    rvr = K.norm_cdf(ph.resampling_normals(nobs, n, cuda_dev))
    u = ph.materialise(nobs, n, cuda_dev)
    a = SP.run_split_smoother(SP.LocalComm(2), obs, params, n, lag, rvr, u_d=u, device=cuda_dev)
    b = SP.run_split_smoother(SP.LocalComm(2), obs, params, n, lag, rvr, philox=(ph.seed, ph.offset),
                              device=cuda_dev)
    torch.cuda.synchronize()
    assert float(a["log_like"].item()) == float(b["log_like"].item())
    assert np.array_equal(a["gradient"].cpu().numpy(), b["gradient"].cpu().numpy())
    # and against the oracle on the materialised numbers
    import oracle
    rvp = np.ascontiguousarray(u.cpu().numpy().T).reshape(-1)
    ref = oracle.flps_sv_corr(obs, params, rvr.cpu().numpy(), rvp, n, lag, 0)
    assert abs(float(a["log_like"].item()) - ref["log_like"]) <= 1e-10 * abs(ref["log_like"])


def test_split_estimator_contract(cuda_dev):
    """SplitParticleMethodsCUDA keeps the estimator contract and agrees with ParticleMethodsCUDA."""
    from pmmh_qn_b200 import DeviceRVS, ParticleMethodsCUDA
    from pmmh_qn_b200.state.particle_methods.split import SplitParticleMethodsCUDA
    from toy_models import ToySVModel
    n, nobs = 3000, 80
    model = ToySVModel(gi.sv_obs(nobs), gi.SV_PARAM_SETS[0])
    rvs = gi.sv_rvs(n, nobs, 7)
    handle = DeviceRVS.from_numpy_particle(rvs, cuda_dev)
    one = ParticleMethodsCUDA(model, no_particles=n, fixed_lag=10)
    assert one.smoother(model, rvs={'rvs': handle})
    sp = SplitParticleMethodsCUDA(model, n, fixed_lag=10, local_world=3)
    assert sp.dim_rvs == one.dim_rvs and sp.alg_type == 'particle'
    assert sp.smoother(model, rvs={'rvs': handle})
    assert abs(sp.results['log_like'] - one.results['log_like']) <= 1e-10 * abs(one.results['log_like'])
    g1, g2 = one.results['gradient_internal'], sp.results['gradient_internal']
    assert np.max(np.abs(g1 - g2)) <= 1e-9 * np.max(np.abs(g1))
    assert relerr(sp.results['smo_state_est'], one.results['smo_state_est']) <= 1e-10
    assert sp.filter(model, rvs={'rvs': handle})
    assert abs(sp.results['log_like'] - one.results['log_like']) <= 1e-10 * abs(one.results['log_like'])


@pytest.mark.parametrize("alg", [4, 5])
@pytest.mark.parametrize("n,nobs,lag,seed", [(20000, 61, 10, 1), (3000, 47, 4, 2), (1 << 17, 45, 10, 4),
                                             (5000, 50, 7, 3), (4000, 40, 2, 5), (4000, 40, 3, 6),
                                             (2500, 70, 13, 7)])
def test_streaming_kernels_behind_flps_sv_corr(cuda_dev, alg, n, nobs, lag, seed):
    """pmmh_flps_sv_corr with algorithm 4 (records) / 5 (path storage + jump tables): the same
    phases driven from C++ on one device; several lags exercise the jump-table decomposition."""
    import torch
    import oracle
    from pmmh_qn_b200 import kernels as K
    obs, params, rvr, rvp = gi.sv_inputs(n, nobs, seed)
    ref = oracle.flps_sv_corr(obs, params, rvr, rvp, n, lag, 0)
    u = torch.from_numpy(to_time_major(rvp, n, nobs)).to(cuda_dev)
    K.set_sv_algorithm(alg)
    try:
        out = K.flps_sv_corr(torch.from_numpy(obs).to(cuda_dev), torch.from_numpy(params).to(cuda_dev),
                             torch.from_numpy(rvr[:nobs].copy()).to(cuda_dev), u, lag=lag)
        torch.cuda.synchronize()
    finally:
        K.set_sv_algorithm(0)
    diag = out["diag"][0].cpu().numpy()
    assert int(diag[6]) == 4 and int(diag[2]) == 0
    ll = float(out["log_like"][0])
    assert abs(ll - ref["log_like"]) <= 1e-10 * abs(ref["log_like"])
    assert relerr(out["filt"][0].cpu().numpy(), ref["filt"]) <= 1e-10
    assert relerr(out["smo"][0].cpu().numpy(), ref["smo"]) <= 1e-10
    assert relerr(out["traj"][0].cpu().numpy(), ref["traj"]) <= 1e-12
    g, gr = out["gradient"][0].cpu().numpy(), ref["gradient"]
    assert np.max(np.abs(g - gr)) <= 1e-9 * np.max(np.abs(gr))


def test_automatic_selection_beyond_the_exchange_kernel(cuda_dev):
    """Default algorithm, N = 1.5 M (the exchange kernel takes at most ~1.16 M on 148 SMs): the
    streaming kernels run and match the oracle."""
    import torch
    import oracle
    from pmmh_qn_b200 import kernels as K
    n, nobs, lag = 1500000, 26, 10
    obs, params, rvr, rvp = gi.sv_inputs(n, nobs, 6)
    ref = oracle.flps_sv_corr(obs, params, rvr, rvp, n, lag, 0)
    u = torch.from_numpy(to_time_major(rvp, n, nobs)).to(cuda_dev)
    out = K.flps_sv_corr(torch.from_numpy(obs).to(cuda_dev), torch.from_numpy(params).to(cuda_dev),
                         torch.from_numpy(rvr[:nobs].copy()).to(cuda_dev), u, lag=lag)
    torch.cuda.synchronize()
    diag = out["diag"][0].cpu().numpy()
    assert int(diag[6]) == 4 and int(diag[2]) == 0
    assert abs(float(out["log_like"][0]) - ref["log_like"]) <= 1e-10 * abs(ref["log_like"])
    g, gr = out["gradient"][0].cpu().numpy(), ref["gradient"]
    assert np.max(np.abs(g - gr)) <= 1e-9 * np.max(np.abs(gr))
    assert relerr(out["smo"][0].cpu().numpy(), ref["smo"]) <= 1e-10


@pytest.mark.parametrize("par", [(0.2, 0.9, 0.4, -0.5), (2.0, 0.9, 0.4, -0.2), (-0.3, 0.97, 0.15, 0.3),
                                 (0.0, 0.995, 0.05, -0.8), (1.0, 0.5, 1.2, 0.0)])
def test_streaming_kernels_agree_with_the_general_kernel_over_parameter_sets(cuda_dev, par):
    """Independent implementations (persistent general kernel, streaming kernels with records and
    with path storage) on the same
    inputs over the parameter sets of the robustness sweep, N = 60 000, T = 300: identical
    near-tie counts are not required, the estimates must agree to the parity tolerances."""
    import torch
    from pmmh_qn_b200 import kernels as K
    n, nobs, lag = 60000, 301, 10
    obs = torch.from_numpy(gi.sv_obs(nobs, params=par)).to(cuda_dev)
    params = torch.tensor([par], dtype=torch.float64, device=cuda_dev)
    g = torch.Generator(device=cuda_dev)
    g.manual_seed(99)
    u = torch.randn((1, nobs, n), dtype=torch.float64, device=cuda_dev, generator=g)
    rvr = torch.rand((1, nobs), dtype=torch.float64, device=cuda_dev, generator=g)
    res = {}
    try:
        for alg in (1, 4, 5):
            K.set_sv_algorithm(alg)
            res[alg] = K.flps_sv_corr(obs, params, rvr, u, lag=lag)
            torch.cuda.synchronize()
    finally:
        K.set_sv_algorithm(0)
    a = res[1]
    assert int(a["diag"][0, 2]) == 0
    for alg in (4, 5):
        b = res[alg]
        assert int(b["diag"][0, 6]) == 4 and int(b["diag"][0, 2]) == 0
        la, lb = float(a["log_like"][0]), float(b["log_like"][0])
        assert abs(la - lb) <= 1e-10 * abs(la), (alg, la, lb)
        ga, gb = a["gradient"][0].cpu().numpy(), b["gradient"][0].cpu().numpy()
        assert np.max(np.abs(ga - gb)) <= 1e-9 * np.max(np.abs(ga)), alg
        assert relerr(b["filt"][0].cpu().numpy(), a["filt"][0].cpu().numpy()) <= 1e-10
        assert relerr(b["smo"][0].cpu().numpy(), a["smo"][0].cpu().numpy()) <= 1e-10
        assert relerr(b["traj"][0].cpu().numpy(), a["traj"][0].cpu().numpy()) <= 1e-12


def test_host_streamed_rvs_on_the_streaming_kernels(cuda_dev):
    """pmmh_flps_sv_corr_streamed beyond the exchange kernel's size (N = 1.5 M): the copy engine fills
    particle-major chunks of 64 time steps while the step loop of the streaming kernels is already running (it waits for a chunk's
    event when it enters it).  Bit-identical to the device-resident evaluation; the staging buffer
    can be reused by back-to-back calls."""
    import torch
    from pmmh_qn_b200 import kernels as K
    n, nobs, lag = 1500000, 130, 10
    dev = cuda_dev
    host = torch.empty((nobs, n + 1), dtype=torch.float64, pin_memory=True)
    g = torch.Generator()
    g.manual_seed(5)
    host.copy_(torch.randn((nobs, n + 1), dtype=torch.float64, generator=g))
    rvs = host.numpy()
    assert K.sv_streamed_eligible(nobs, n, lag)
    obs = torch.from_numpy(gi.sv_obs(nobs)).to(dev)
    params = torch.tensor([[0.2, 0.9, 0.4, -0.5]], dtype=torch.float64, device=dev)
    rvr_h, rvp = gi.split_particle(rvs, nobs)      # (Phi of the first n_obs flat entries, rvp)
    rvr = torch.from_numpy(rvr_h).to(dev)
    ws, st = K.Workspace(), K.Workspace()
    a = K.flps_sv_corr_streamed(rvs, obs, params, rvr, nobs, n, lag=lag, workspace=ws, stage=st)
    a2 = K.flps_sv_corr_streamed(rvs, obs, params, rvr, nobs, n, lag=lag, workspace=ws, stage=st)
    torch.cuda.synchronize()
    assert int(a["diag"][0, 6]) == 4 and int(a["diag"][0, 2]) == 0
    u = torch.from_numpy(to_time_major(rvp, n, nobs)).to(dev)
    b = K.flps_sv_corr(obs, params, rvr, u, lag=lag, compute_hessian=False)
    torch.cuda.synchronize()
    assert int(b["diag"][0, 6]) == 4
    for k in ("log_like", "filt", "smo", "gradient", "traj"):
        assert torch.equal(a[k], b[k]), k
        assert torch.equal(a2[k], b[k]), k


def test_philox_entry_point_on_one_device(cuda_dev):
    """pmmh_flps_sv_corr_philox (path storage, u regenerated from the stream) == the same stream
    materialised and passed to pmmh_flps_sv_corr, and == the split filter's Philox mode."""
    import torch
    from pmmh_qn_b200 import kernels as K
    from pmmh_qn_b200.state.particle_methods import split as SP
    n, nobs, lag = 40000, 70, 10
    obs = torch.from_numpy(gi.sv_obs(nobs)).to(cuda_dev)
    params = torch.tensor([gi.SV_PARAM_SETS[0]], dtype=torch.float64, device=cuda_dev)
    ph = SP.PhiloxRVS(seed=77, offset=5)
    rvr = K.norm_cdf(ph.resampling_normals(nobs, n, cuda_dev))
    a = K.flps_sv_corr_philox(obs, params, rvr, ph.seed, ph.offset, n, lag=lag)
    u = ph.materialise(nobs, n, cuda_dev)
    K.set_sv_algorithm(5)
    try:
        b = K.flps_sv_corr(obs, params, rvr, u, lag=lag)
        torch.cuda.synchronize()
    finally:
        K.set_sv_algorithm(0)
    assert int(a["diag"][0, 6]) == 4 and int(a["diag"][0, 2]) == 0
    for k in ("log_like", "filt", "smo", "gradient", "traj"):
        assert torch.equal(a[k], b[k]), k
    c = SP.run_split_smoother(SP.LocalComm(2), gi.sv_obs(nobs), np.array(gi.SV_PARAM_SETS[0]), n, lag, rvr,
                              philox=(ph.seed, ph.offset), device=cuda_dev)
    la, lc = float(a["log_like"][0]), float(c["log_like"].item())
    assert abs(la - lc) <= 1e-10 * abs(la)
