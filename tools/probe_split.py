"""Timing probe of the split particle filter on ONE GPU (LocalComm): python tools/probe_split.py N T world lag"""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
import numpy as np
import torch

import golden_inputs as gi
from pmmh_qn_b200 import kernels as K
from pmmh_qn_b200.state.particle_methods import split as SP

n, T, world, lag = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4])
dev = torch.device("cuda:0")
nobs = T + 1
obs = gi.sv_obs(nobs)
params = np.array(gi.SV_PARAM_SETS[0], dtype=np.float64)
ph = SP.PhiloxRVS(seed=5, offset=0)
rvr = K.norm_cdf(ph.resampling_normals(nobs, n, dev))
for rep in range(2):
    torch.cuda.synchronize()
    t0 = time.time()
    out = SP.run_split_smoother(SP.LocalComm(world), obs, params, n, lag, rvr, philox=(5, 0), device=dev)
    torch.cuda.synchronize()
    dt = time.time() - t0
    print(json.dumps({"N": n, "T": T, "world_in_process": world, "lag": lag, "seconds": dt,
                      "particle_steps_per_s": n * T / dt, "log_like": float(out["log_like"].item()),
                      "near_ties": sum(o["diag"][0] for o in out["per_rank"]),
                      "max_bin": max(o["diag"][1] for o in out["per_rank"]),
                      "max_arrivals": int(out["counts"].max()),
                      "children_max_over_mean": float((out["children"].max(axis=1) / (n / world)).mean()),
                      "children_worst_step": float(out["children"].max() / (n / world))}), flush=True)
