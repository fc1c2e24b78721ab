#!/usr/bin/env python
"""Development probe: run the exchange kernel without fallback and print the failure dump."""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tools")); sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
import numpy as np, torch
import golden_inputs as gi
from pmmh_qn_b200 import kernels as K
K.set_sv_algorithm(2)
n = int(sys.argv[1]); nobs = int(sys.argv[2])
dev = torch.device("cuda:0")
g = torch.Generator(device=dev); g.manual_seed(1234)
obs = torch.from_numpy(gi.sv_obs(nobs)).to(dev)
params = torch.tensor([[0.2, 0.9, 0.4, -0.5]], dtype=torch.float64, device=dev)
u = torch.randn((1, nobs, n), dtype=torch.float64, device=dev, generator=g)
rvr = torch.rand((1, nobs), dtype=torch.float64, device=dev, generator=g)
out = K.flps_sv_corr(obs, params, rvr, u, lag=10, compute_hessian=False)
torch.cuda.synchronize()
d = out["diag"][0].tolist()
print("diag", d, "reason", d[7] & 255, "step", (d[7] >> 8) & 0xffffff, "maxarr", d[7] >> 32)
print("ll", float(out["log_like"][0]))
np.set_printoptions(linewidth=200)
print("hess1", out["hess1"][0].cpu().numpy().reshape(-1))
print("hess2", out["hess2"][0].cpu().numpy().reshape(-1))
