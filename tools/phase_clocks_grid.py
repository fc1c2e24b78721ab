#!/usr/bin/env python
"""Per-CTA phase clocks of the grid kernel (sv_grid.cu): usage phase_clocks_grid.py logN [T] [ctas]."""
import ctypes
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
import numpy as np
import torch

import golden_inputs as gi
from pmmh_qn_b200 import kernels as K, _lib

logn = int(sys.argv[1])
T = int(sys.argv[2]) if len(sys.argv) > 2 else 1000
ctas = int(sys.argv[3]) if len(sys.argv) > 3 else 0
HESS = bool(os.environ.get("PMMH_PROBE_HESS"))   # the Hessian branch
dev = torch.device("cuda:0")
n, nobs = 1 << logn, T + 1
obs = torch.from_numpy(gi.sv_obs(nobs)).to(dev)
params = torch.tensor(gi.SV_PARAM_SETS[0], dtype=torch.float64, device=dev)
g = torch.Generator(device=dev)
g.manual_seed(0)
u = torch.randn((nobs, n), dtype=torch.float64, device=dev, generator=g)
rvr = torch.rand((nobs,), dtype=torch.float64, device=dev, generator=g)
K.set_sv_algorithm(6)
ws = K.Workspace()
out = K.flps_sv_corr(obs, params, rvr, u, lag=10, workspace=ws, ctas_per_problem=ctas, compute_hessian=HESS)
torch.cuda.synchronize()
buf = torch.zeros((160, 32), dtype=torch.int64, device=dev)
_lib.load().pmmh_sv_debug_profile(ctypes.c_void_p(buf.data_ptr()))
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
out = K.flps_sv_corr(obs, params, rvr, u, lag=10, workspace=ws, ctas_per_problem=ctas, compute_hessian=HESS)
e1.record()
torch.cuda.synchronize()
_lib.load().pmmh_sv_debug_profile(None)
ms = e0.elapsed_time(e1)
d = out["diag"][0].cpu().numpy()
c = buf.cpu().numpy().astype(np.float64)
c = c[c.sum(axis=1) > 0]
names = ["C:ranges+fill", "zero hist", "wait 4", "A1:children+hist", "A1:records", "wait 1", "A2:scan+scatter",
         "zero+prefetch", "wait 2", "B:bin sort", "B:rank+wts+score", "B:scan+publish", "wait 3", "score terms", "alpha (Hessian)", "-",
         "C1:totals->offsets", "C2:cumsum+counts", "C3:running max", "A1a:gather+propagate", "A2a:hist scan+tile map",
         "A2b:slot counting", "A2c:reservation", "B1:mailbox+subbin count", "B2:sub-bin scan", "B4:exact ranks", "B5:reorder"]
# (slots 16+ split the phases above: C = C1 + C2 + C3 + [0], A1 = A1a + [3], A2 = A2a + A2b + A2c + [6],
#  bin sort = B1 + B2 + [9], rank+wts = B4 + B5 + [10])
clk = c.sum(axis=1).mean() / (ms * 1e-3) / 1e6   # MHz seen by clock64
print(json.dumps({"N": n, "T": T, "ms": ms, "us_per_step": ms * 1e3 / T, "particle_steps_per_s": n * T / ms * 1e3,
                  "kernel": int(d[6]), "status": int(d[2]), "info": int(d[7]), "near_ties": int(d[0]),
                  "max_bin": int(d[1]), "ctas": int(c.shape[0]), "clock_mhz": clk,
                  "log_like": float(out["log_like"][0])}))
print("per-step microseconds (mean / min / max over CTAs), instrumented run")
for k, nm in enumerate(names):
    v = c[:, k] / T / clk
    print("  %-24s %7.2f %7.2f %7.2f" % (nm, v.mean(), v.min(), v.max()))
