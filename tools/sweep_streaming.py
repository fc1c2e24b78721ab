#!/usr/bin/env python
"""Robustness sweep of the streaming kernels at a size beyond the exchange kernel: parameter sets x
seeds at N = 2^21, T = 300; path storage (algorithm 5, the automatic choice) against the record
variant (algorithm 4).  Prints one JSON line per run."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
import numpy as np
import torch

import golden_inputs as gi
from pmmh_qn_b200 import kernels as K

dev = torch.device("cuda:0")
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 21
nobs = 301
ws = K.Workspace()
for pi, par in enumerate(gi.SV_PARAM_SETS + [(0.0, 0.995, 0.05, -0.8), (1.0, 0.5, 1.2, 0.0)]):
    obs = torch.from_numpy(gi.sv_obs(nobs, params=par)).to(dev)
    params = torch.tensor([par], dtype=torch.float64, device=dev)
    for seed in range(2):
        g = torch.Generator(device=dev)
        g.manual_seed(4321 + seed)
        u = torch.randn((1, nobs, n), dtype=torch.float64, device=dev, generator=g)
        rvr = torch.rand((1, nobs), dtype=torch.float64, device=dev, generator=g)
        res = {}
        for alg in (0, 4):
            K.set_sv_algorithm(alg)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            out = K.flps_sv_corr(obs, params, rvr, u, lag=10, workspace=ws)
            e1.record()
            torch.cuda.synchronize()
            res[alg] = (out, e0.elapsed_time(e1))
        K.set_sv_algorithm(0)
        a, b = res[0][0], res[4][0]
        ga, gb = a["gradient"][0].sum(dim=1), b["gradient"][0].sum(dim=1)
        d = a["diag"][0].tolist()
        print(json.dumps(dict(params=par, seed=seed, N=n, T=nobs - 1, ms_path=round(res[0][1], 2),
                              ms_records=round(res[4][1], 2), kernel=d[6], status=d[2], near_ties=d[0], max_bin=d[1],
                              ll=float(a["log_like"][0]),
                              ll_rel_diff=abs(float(a["log_like"][0]) - float(b["log_like"][0])) / abs(float(b["log_like"][0])),
                              grad_rel_diff=float((ga - gb).abs().max() / gb.abs().max()))), flush=True)
        del u
