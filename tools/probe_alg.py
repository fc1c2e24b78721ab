"""Time pmmh_flps_sv_corr for one algorithm: python tools/probe_alg.py ALG logN T [reps]  (PMMH_PROBE_HESS=1: with the Hessian branch)"""
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

alg, logn, T = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
reps = int(sys.argv[4]) if len(sys.argv) > 4 else 3
dev = torch.device("cuda:0")
n, nobs = 1 << logn, T + 1
obs = torch.from_numpy(gi.sv_obs(nobs)).to(dev)
params = torch.tensor(gi.SV_PARAM_SETS[0], dtype=torch.float64, device=dev)
g = torch.Generator(device=dev)
g.manual_seed(0)
u = torch.randn((nobs, n), dtype=torch.float64, device=dev, generator=g)
rvr = torch.rand((nobs,), dtype=torch.float64, device=dev, generator=g)
K.set_sv_algorithm(alg)
ws = K.Workspace()
stream = torch.cuda.Stream() if os.environ.get("PMMH_SPLIT_GRAPH") else torch.cuda.current_stream()
torch.cuda.synchronize()
torch.cuda.set_stream(stream)
for rep in range(reps):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    out = K.flps_sv_corr(obs, params, rvr, u, lag=10, workspace=ws, compute_hessian=bool(os.environ.get("PMMH_PROBE_HESS")))
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    d = out["diag"][0].cpu().numpy()
    print(json.dumps({"alg": alg, "N": n, "T": T, "ms": ms, "particle_steps_per_s": n * T / ms * 1e3,
                      "log_like": float(out["log_like"][0]), "kernel": int(d[6]), "status": int(d[2]),
                      "near_ties": int(d[0]), "grad0": float(out["gradient"][0].sum(dim=1)[0]),
                      "hess1_00": float(out["hess1"][0].flatten()[0]), "hess2_23": float(out["hess2"][0].flatten()[11])}), flush=True)
