#!/usr/bin/env python
"""Robustness sweep of the exchange kernel (no fallback): seeds x parameter sets at N = 2^20, T = 1000.
Prints status / abandon reason, max arrivals per CTA and the time of every run."""
import os, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
import numpy as np, torch
import golden_inputs as gi
from pmmh_qn_b200 import kernels as K
K.set_sv_algorithm(2)
dev = torch.device("cuda:0")
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 20
nobs = 1001
ws = K.Workspace()
for pi, par in enumerate(gi.SV_PARAM_SETS + [(0.0, 0.995, 0.05, -0.8), (1.0, 0.5, 1.2, 0.0)]):
    obs = torch.from_numpy(gi.sv_obs(nobs, params=par if abs(par[1]) < 1 else gi.SV_PARAM_SETS[0])).to(dev)
    params = torch.tensor([par], dtype=torch.float64, device=dev)
    for seed in range(4 if pi == 0 else 2):
        g = torch.Generator(device=dev); g.manual_seed(1234 + seed)
        u = torch.randn((1, nobs, n), dtype=torch.float64, device=dev, generator=g)
        rvr = torch.rand((1, nobs), dtype=torch.float64, device=dev, generator=g)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        out = K.flps_sv_corr(obs, params, rvr, u, lag=10, compute_hessian=False, workspace=ws)
        e1.record(); torch.cuda.synchronize()
        d = out["diag"][0].tolist()
        print(json.dumps(dict(params=par, seed=seed, ms=round(e0.elapsed_time(e1), 2), status=d[2], reason=d[7] & 255,
                              step=(d[7] >> 8) & 0xffffff, max_arrivals=d[7] >> 32, max_chunk=d[1], near_ties=d[0],
                              ll=float(out["log_like"][0]))), flush=True)
        del u
