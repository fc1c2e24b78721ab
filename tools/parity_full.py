#!/usr/bin/env python
"""Parity at the headline shape itself: N = 2^20, T = 1000 on the bench inputs, CUDA path (automatic
kernel selection = the grid kernel) against the CPU oracle; ancestors compared at every time step.
Takes ~10 minutes of one host core (the oracle is sequential).  usage: parity_full.py [logN] [T]"""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests", "golden")):
    sys.path.insert(0, p)
import numpy as np
import torch

import golden_inputs as gi
import oracle
from pmmh_qn_b200 import kernels as K

logn = int(sys.argv[1]) if len(sys.argv) > 1 else 20
T = int(sys.argv[2]) if len(sys.argv) > 2 else 1000
n, nobs, lag = 1 << logn, T + 1, 10
dev = torch.device("cuda:0")
params_h = np.array([0.2, 0.9, 0.4, -0.5])
obs_h = gi.sv_obs(nobs)
g = torch.Generator(device=dev)
g.manual_seed(1234)          # bench.py rank 0
u = torch.randn((1, nobs, n), dtype=torch.float64, device=dev, generator=g)
rvr = torch.rand((1, nobs), dtype=torch.float64, device=dev, generator=g)
out = K.flps_sv_corr(torch.from_numpy(obs_h).to(dev), torch.tensor([params_h], dtype=torch.float64, device=dev),
                     rvr, u, lag=lag, compute_hessian=False, store_history=True)
torch.cuda.synchronize()
d = out["diag"][0].tolist()
A = out["A"][0].cpu().numpy()
X = out["X"][0].cpu().numpy()
res = {k: out[k][0].cpu().numpy() for k in ("filt", "smo", "gradient", "traj")}
ll = float(out["log_like"][0])
del out
rvp = np.empty(n * nobs)
uh = u[0].cpu().numpy()
rvp.reshape(n, nobs)[:] = uh.T
del uh, u
t0 = time.perf_counter()
ref = oracle.flps_sv_corr(obs_h, params_h, rvr[0].cpu().numpy(), rvp, n, lag, 0, dumps=True)
secs = time.perf_counter() - t0
neq = np.any(A[1:] != ref["A"][1:], axis=1)
per_step = np.sum(A[1:] != ref["A"][1:], axis=1)
xdiff = np.sum(X[1:] != ref["X"][1:], axis=1)
rel = lambda a, b: float(np.max(np.abs(a - b)) / max(1e-300, np.max(np.abs(b))))
print(json.dumps({
    "what": "CUDA path vs oracle at the headline shape", "N": n, "T": T, "lag": lag, "kernel": int(d[6]),
    "status": int(d[2]), "near_ties": int(d[0]), "soft_ties": int(d[4]), "generations_mismatched": int(neq.sum()),
    "first_mismatch_step": (int(np.argmax(neq)) + 1) if neq.any() else None,
    "x_max_rel": rel(X, ref["X"]), "ll": ll, "ll_oracle": ref["log_like"],
    "ll_rel": abs(ll - ref["log_like"]) / abs(ref["log_like"]),
    "grad_rel": rel(res["gradient"], np.asarray(ref["gradient"])), "filt_rel": rel(res["filt"], ref["filt"]),
    "smo_rel": rel(res["smo"], ref["smo"]), "traj_rel": rel(res["traj"], ref["traj"]),
    "ancestor_entries_differing_max_per_step": int(per_step.max()), "ancestor_entries_differing_total": int(per_step.sum()),
    "particle_values_differing_max_per_step": int(xdiff.max()), "steps_with_differing_values": int((xdiff > 0).sum()),
    "episodes": int(np.sum((xdiff[1:] > 0) & (xdiff[:-1] == 0)) + (xdiff[0] > 0)),
    "oracle_seconds": secs}))
