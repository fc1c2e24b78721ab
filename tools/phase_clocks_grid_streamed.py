#!/usr/bin/env python
"""Per-CTA phase clocks of the grid kernel with HOST-resident u (pmmh_flps_sv_corr_streamed): usage [logN] [T]."""
import ctypes, json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
import numpy as np, torch
import golden_inputs as gi
from pmmh_qn_b200 import kernels as K, _lib
logn = int(sys.argv[1]) if len(sys.argv) > 1 else 20
T = int(sys.argv[2]) if len(sys.argv) > 2 else 1000
dev = torch.device("cuda:0")
n, nobs = 1 << logn, T + 1
host = torch.empty((nobs, n + 1), dtype=torch.float64, pin_memory=True)
g = torch.Generator(device=dev); g.manual_seed(0)
for r0 in range(0, nobs, 64):
    host[r0:r0 + 64].copy_(torch.randn((min(64, nobs - r0), n + 1), dtype=torch.float64, device=dev, generator=g))
rvs = host.numpy()
obs = torch.from_numpy(gi.sv_obs(nobs)).to(dev)
params = torch.tensor([gi.SV_PARAM_SETS[0]], dtype=torch.float64, device=dev)
rvr = torch.rand((nobs,), dtype=torch.float64, device=dev, generator=g)
ws, st = K.Workspace(), K.Workspace()
out = K.flps_sv_corr_streamed(rvs, obs, params, rvr, nobs, n, lag=10, workspace=ws, stage=st)
torch.cuda.synchronize()
buf = torch.zeros((160, 32), dtype=torch.int64, device=dev)
_lib.load().pmmh_sv_debug_profile(ctypes.c_void_p(buf.data_ptr()))
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
out = K.flps_sv_corr_streamed(rvs, obs, params, rvr, nobs, n, lag=10, workspace=ws, stage=st)
e1.record(); torch.cuda.synchronize()
_lib.load().pmmh_sv_debug_profile(None)
ms = e0.elapsed_time(e1)
c = buf.cpu().numpy().astype(np.float64); c = c[c.sum(axis=1) > 0]
names = ["C:ranges+fill", "zero hist", "wait 4 (+ data)", "A1:children+hist", "A1:records", "wait 1", "A2:scan+scatter",
         "zero+prefetch", "wait 2", "B:bin sort", "B:rank+wts+score", "B:scan+publish", "wait 3", "score terms", "-", "-",
         "C1", "C2", "C3", "A1a:gather+propagate", "A2a", "A2b", "A2c", "B1", "B2", "B4", "B5"]
clk = 1.92e3   # MHz (clock64), nominal
print(json.dumps({"N": n, "T": T, "ms": ms, "status": int(out["diag"][0, 2])}))
tot = 0.0
for k, nm in enumerate(names):
    v = c[:, k] / T / clk
    if nm != "-": print("  %-24s %7.2f %7.2f %7.2f" % (nm, v.mean(), v.min(), v.max()))
    tot += v.mean()
print("  sum %.2f us per step; without 'wait 4 (+ data)': %.2f" % (tot, tot - (c[:, 2] / T / clk).mean()))
