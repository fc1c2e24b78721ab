#!/usr/bin/env python
"""Secondary measurements (one GPU): the other BASELINE configs and the auxiliary kernels, each
with its HBM roofline fraction (algorithmic bytes per unit from SURVEY.md section 8d).

  config 1  random-effects importance sampler, 100 x 100, batches of 1 / 1024 / 65536 evaluations
  config 3  data-subsampling estimator, n = 11 M x 28 (2.46 GB in HBM), m = 550 000:
            Crank-Nicolson + Phi + sort + stratified indices + gather-reduce (grad, grad + Hessian)
  config 4  1024 SV chains x N = 4096, T = 1000 (one CTA per chain)
  aux       Crank-Nicolson over 2^29 doubles, the rvs layout change (transpose) at T=1000 N=2^20

Prints one JSON object per line; run under gpurun, results are copied into profiles/ by hand.
"""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
import golden_inputs as gi  # noqa: E402
from pmmh_qn_b200 import kernels as K  # noqa: E402

PEAK = 6515.7e9
try:
    PEAK = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]) * 1e9
except Exception:
    pass
dev = torch.device("cuda:0")


def timed(fn, reps=5, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e-3)
    return float(np.median(ts))


def emit(**kw):
    print(json.dumps(kw), flush=True)


def bench_importance():
    nobs = n = 100
    obs_r, par_r, rvr_r, rvp_r = gi.re_inputs(n, nobs, 0)
    obs = torch.from_numpy(obs_r).to(dev)
    for B in (1, 1024, 65536):
        g = torch.Generator(device=dev)
        g.manual_seed(B)
        params = torch.tensor(par_r, dtype=torch.float64, device=dev).repeat(B, 1).contiguous()
        rvr = torch.rand((B,), dtype=torch.float64, device=dev, generator=g)
        rvp = torch.randn((B, nobs * n), dtype=torch.float64, device=dev, generator=g)
        t = timed(lambda: K.importance_discrete(obs, params, rvr, rvp, nobs, n))
        byt = B * nobs * n * 8
        emit(what="config1 importance_discrete 100x100", batch=B, seconds=t, evals_per_s=B / t,
             achieved_gbs=byt / t / 1e9, roofline_frac=byt / t / PEAK)


def bench_subsampling():
    n, d, m = 11_000_000, 28, 550_000
    g = torch.Generator(device=dev)
    g.manual_seed(0)
    x = torch.randn((n, d), dtype=torch.float64, device=dev, generator=g)
    beta = 0.1 * torch.randn((d,), dtype=torch.float64, device=dev, generator=g)
    y = (torch.rand((n,), dtype=torch.float64, device=dev, generator=g) <
         torch.sigmoid(x @ beta)).to(torch.float64)
    u = torch.randn((m,), dtype=torch.float64, device=dev, generator=g)
    ws1, ws2 = K.Workspace(), K.Workspace()
    t_cn = timed(lambda: K.crank_nicolson(u, 0.05, seed=1))
    t_idx = timed(lambda: K.subsample_indices(u, n, workspace=ws1))
    idx = K.subsample_indices(u, n, workspace=ws1)
    for hess in (False, True):
        t_red = timed(lambda: K.logistic_loglike(x, y, idx, beta, compute_hessian=hess, workspace=ws2))
        byt = m * (8 * d + 12)
        emit(what="config3 subsampling n=11M d=28 m=550000", hessian=int(hess), cn_seconds=t_cn,
             phi_sort_index_seconds=t_idx, gather_reduce_seconds=t_red,
             rows_per_s=m / (t_cn + t_idx + t_red), gather_reduce_gbs=byt / t_red / 1e9,
             gather_reduce_roofline_frac=byt / t_red / PEAK)
    del x, y


def bench_chains():
    n, nobs, B = 4096, 1001, 1024
    g = torch.Generator(device=dev)
    g.manual_seed(4)
    obs = torch.from_numpy(gi.sv_obs(nobs)).to(dev)
    base = torch.tensor([0.2, 0.9, 0.4, -0.5], dtype=torch.float64, device=dev)
    params = (base + 0.01 * torch.randn((B, 4), dtype=torch.float64, device=dev, generator=g)).contiguous()
    u = torch.randn((B, nobs, n), dtype=torch.float64, device=dev, generator=g)
    rvr = torch.rand((B, nobs), dtype=torch.float64, device=dev, generator=g)
    ws = K.Workspace()
    for hess in (False, True):
        t = timed(lambda: K.flps_sv_corr(obs, params, rvr, u, lag=10, compute_hessian=hess, workspace=ws),
                  reps=3, warm=1)
        steps = B * n * (nobs - 1)
        byt = steps * (192 if hess else 96)
        emit(what="config4 1024 SV chains x N=4096 T=1000", hessian=int(hess), seconds=t,
             particle_steps_per_s=steps / t, loglik_evals_per_s=B / t, roofline_frac=byt / t / PEAK)
    del u


def bench_elementwise():
    nel = 1 << 29
    g = torch.Generator(device=dev)
    g.manual_seed(9)
    u = torch.randn((nel,), dtype=torch.float64, device=dev, generator=g)
    out = torch.empty_like(u)
    t = timed(lambda: K.crank_nicolson(u, 0.05, seed=3, out=out))
    emit(what="crank_nicolson philox 2^29 doubles", seconds=t, achieved_gbs=16 * nel / t / 1e9,
         roofline_frac=16 * nel / t / PEAK)
    xi = torch.randn((nel,), dtype=torch.float64, device=dev, generator=g)
    t = timed(lambda: K.crank_nicolson(u, 0.05, xi=xi, out=out))
    emit(what="crank_nicolson supplied xi 2^29 doubles", seconds=t, achieved_gbs=24 * nel / t / 1e9,
         roofline_frac=24 * nel / t / PEAK)
    del xi, out, u
    nobs, n = 1001, 1 << 20
    rvs = torch.randn((nobs * (n + 1),), dtype=torch.float64, device=dev, generator=g)
    t = timed(lambda: K.split_rvs(rvs, nobs, n), reps=3, warm=1)
    emit(what="split_rvs (layout change) T=1000 N=2^20", seconds=t, achieved_gbs=16 * nobs * n / t / 1e9,
         roofline_frac=16 * nobs * n / t / PEAK)


if __name__ == "__main__":
    which = sys.argv[1:] or ["importance", "elementwise", "subsampling", "chains"]
    for w in which:
        {"importance": bench_importance, "subsampling": bench_subsampling, "chains": bench_chains,
         "elementwise": bench_elementwise}[w]()
        torch.cuda.empty_cache()
