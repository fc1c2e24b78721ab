"""Small shapes of the SV kernels for compute-sanitizer (memcheck / racecheck / synccheck):
   compute-sanitizer --tool racecheck python tools/sanitize_small.py [alg ...]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import torch

import golden_inputs as gi
from helpers import to_time_major
from pmmh_qn_b200 import kernels as K

algs = [int(a) for a in sys.argv[1:]] or [6]
dev = torch.device("cuda:0")
for alg in algs:
    n, nobs, ctas = (3000, 24, 1) if alg == 3 else (6000, 24, 4)
    obs, params, rvr, rvp = gi.sv_inputs(n, nobs, 1)
    u = to_time_major(rvp, n, nobs)
    K.set_sv_algorithm(alg)
    out = K.flps_sv_corr(torch.from_numpy(obs).to(dev), torch.from_numpy(params).to(dev),
                         torch.from_numpy(rvr[:nobs].copy()).to(dev), torch.from_numpy(u).to(dev), lag=10,
                         compute_hessian=False, ctas_per_problem=(0 if alg in (3, 4, 5) else ctas))
    torch.cuda.synchronize()
    d = out["diag"][0].cpu().numpy()
    print("alg %d: log_like %.12f kernel %d status %d" % (alg, float(out["log_like"][0]), int(d[6]), int(d[2])), flush=True)
K.set_sv_algorithm(0)
if "streamed" in os.environ.get("PMMH_SANITIZE_EXTRA", "streamed"):
    # host-resident rvs: the copy engine feeds the running kernel (flag polled per time step)
    for alg, n, nobs in ((6, 6000, 40), (2, 6000, 40)):
        rvs = gi.sv_rvs(n, nobs, 2)
        rvr_h, rvp = gi.split_particle(rvs, nobs)
        obs = gi.sv_obs(nobs)
        K.set_sv_algorithm(alg)
        out = K.flps_sv_corr_streamed(np.ascontiguousarray(rvs), torch.from_numpy(obs).to(dev),
                                      torch.tensor([gi.SV_PARAM_SETS[0]], dtype=torch.float64, device=dev),
                                      torch.from_numpy(rvr_h).to(dev), nobs, n, lag=10,
                                      ctas_per_problem=(0 if alg == 6 else 4))
        torch.cuda.synchronize()
        d = out["diag"][0].cpu().numpy()
        print("streamed alg %d: log_like %.12f kernel %d status %d" % (alg, float(out["log_like"][0]), int(d[6]), int(d[2])), flush=True)
    K.set_sv_algorithm(0)
