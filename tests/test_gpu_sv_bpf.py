"""GPU parity of pmmh_bpf_sv_corr (bootstrap filter) against the oracle / goldens.

PMMH_BPF_PARITY reproduces the reference's Q2 behaviour (leverage term reads the time-i column,
a serial dependency chain resolved on the device as a wavefront); PMMH_BPF_INTENDED is the
evidently intended time i-1 read and is checked against the oracle's intended_read switch."""
import numpy as np
import pytest

import golden_inputs as gi
from helpers import first_mismatch_step, relerr, to_time_major

pytestmark = pytest.mark.gpu


def _run(dev, obs, params, rvr, rvp, n, nobs, mode, hist=True, ctas=0):
    import torch
    from pmmh_qn_b200 import kernels as K
    out = K.bpf_sv_corr(torch.from_numpy(obs).to(dev), torch.from_numpy(params).to(dev),
                        torch.from_numpy(rvr[:nobs].copy()).to(dev),
                        torch.from_numpy(to_time_major(rvp, n, nobs)).to(dev), read_mode=mode,
                        store_history=hist, ctas_per_problem=ctas)
    torch.cuda.synchronize()
    return {k: v.cpu().numpy() for k, v in out.items() if not k.startswith("_")}


def _check(res, ref):
    if not np.isfinite(ref["log_like"]):
        assert not np.isfinite(res["log_like"][0])
        return
    step = first_mismatch_step(res["A"][0][1:], ref["A"][1:])
    assert step is None, "ancestors differ first at time %d" % (step + 1)
    assert relerr(res["X"][0], ref["X"]) <= 1e-12
    assert abs(res["log_like"][0] - ref["log_like"]) <= 1e-10 * abs(ref["log_like"])
    assert relerr(res["filt"][0], ref["filt"]) <= 1e-10
    assert relerr(res["traj"][0], ref["traj"]) <= 1e-12
    assert int(res["diag"][0, 5]) == ref["traj_idx"]


@pytest.mark.parametrize("n,nobs", [(75, 361), (37, 50), (200, 120), (1024, 301)])
@pytest.mark.parametrize("seed", [0, 1])
@pytest.mark.parametrize("mode", [0, 1])
def test_bpf_vs_oracle(cuda_dev, n, nobs, seed, mode):
    import oracle
    obs, params, rvr, rvp = gi.sv_inputs(n, nobs, seed)
    ref = oracle.bpf_sv_corr(obs, params, rvr, rvp, n, intended_read=bool(mode), dumps=True)
    res = _run(cuda_dev, obs, params, rvr, rvp, n, nobs, mode)
    _check(res, ref)


@pytest.mark.parametrize("ctas", [1, 3, 148])
def test_bpf_multi_cta(cuda_dev, ctas):
    import oracle
    n, nobs = 3000, 120
    obs, params, rvr, rvp = gi.sv_inputs(n, nobs, 0)
    for mode in (0, 1):
        ref = oracle.bpf_sv_corr(obs, params, rvr, rvp, n, intended_read=bool(mode), dumps=True)
        res = _run(cuda_dev, obs, params, rvr, rvp, n, nobs, mode, ctas=ctas)
        _check(res, ref)


def test_bpf_vs_golden(cuda_dev, golden):
    g = golden["sv_kernels"]
    for (n, nobs, lag, seeds) in gi.SV_KERNEL_CASES:
        if n > 1024:
            continue
        for seed in seeds:
            obs, params, rvr, rvp = gi.sv_inputs(n, nobs, seed)
            pre = "bpf_n%d_t%d_l%d_s%d_" % (n, nobs, lag, seed)
            ll = float(g[pre + "log_like"])
            res = _run(cuda_dev, obs, params, rvr, rvp, n, nobs, 0, hist=False)
            if not np.isfinite(ll):
                assert not np.isfinite(res["log_like"][0])
                continue
            assert abs(res["log_like"][0] - ll) <= 1e-10 * abs(ll)
            assert relerr(res["filt"][0], g[pre + "filt"]) <= 1e-10
            assert relerr(res["traj"][0], g[pre + "traj"]) <= 1e-12
