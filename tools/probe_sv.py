#!/usr/bin/env python
"""Quick timing probe of the SV smoother kernel (not the bench; used while developing)."""
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
import golden_inputs as gi  # noqa: E402
from pmmh_qn_b200 import kernels as K  # noqa: E402


def run(n, nobs=1001, batch=1, hess=False, hist=False, ctas=0, reps=3, lag=10):
    dev = torch.device("cuda:0")
    g = torch.Generator(device=dev)
    g.manual_seed(1234)
    obs = torch.from_numpy(gi.sv_obs(nobs)).to(dev)
    params = torch.tensor([[0.2, 0.9, 0.4, -0.5]] * batch, dtype=torch.float64, device=dev)
    u = torch.randn((batch, nobs, n), dtype=torch.float64, device=dev, generator=g)
    rvr = torch.rand((batch, nobs), dtype=torch.float64, device=dev, generator=g)
    ws = K.Workspace()
    times = []
    out = None
    for r in range(reps + 1):
        e0 = torch.cuda.Event(enable_timing=True)
        e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
        out = K.flps_sv_corr(obs, params, rvr, u, lag=lag, compute_hessian=hess, store_history=hist,
                             ctas_per_problem=ctas, workspace=ws)
        e1.record()
        torch.cuda.synchronize()
        if r > 0:
            times.append(e0.elapsed_time(e1))
    ms = float(np.median(times))
    steps = batch * n * (nobs - 1)
    bytes_per = 192 if hess else 96
    rec = dict(n=n, batch=batch, hess=int(hess), hist=int(hist), ctas=ctas, ms=round(ms, 3),
               particle_steps_per_s=steps / (ms * 1e-3),
               roofline_frac=steps * bytes_per / (ms * 1e-3) / 6515.7e9,
               ll=float(out["log_like"][0]), diag=out["diag"][0].tolist())
    print(json.dumps(rec), flush=True)
    return rec


if __name__ == "__main__":
    which = sys.argv[1] if len(sys.argv) > 1 else "all"
    if which == "one":
        run(int(sys.argv[2]), hess=bool(int(sys.argv[3])), reps=int(sys.argv[4]))
    if which in ("all", "single"):
        for n in (4096, 65536, 262144, 1048576):
            run(n)
        run(1048576, hist=True)
        run(1048576, hess=True)
    if which in ("all", "batch"):
        run(4096, batch=148)
        run(4096, batch=1024, reps=2)
