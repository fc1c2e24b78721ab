#!/usr/bin/env python
"""Per-CTA phase clocks of the chain kernel: usage phase_clocks_chain.py N [NOBS]."""
import sys, os, ctypes
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tools"))
import numpy as np, torch
import probe_sv
from pmmh_qn_b200 import kernels as K, _lib
K.set_sv_algorithm(3)
n = int(sys.argv[1]); nobs = int(sys.argv[2]) if len(sys.argv) > 2 else 1001
probe_sv.run(n, nobs=nobs, reps=1)
buf = torch.zeros((148, 16), dtype=torch.int64, device="cuda:0")
_lib.load().pmmh_sv_debug_profile(ctypes.c_void_p(buf.data_ptr()))
probe_sv.run(n, nobs=nobs, reps=0)
torch.cuda.synchronize()
_lib.load().pmmh_sv_debug_profile(None)
c = buf.cpu().numpy().astype(np.float64)[0]
names = ["resample+propagate", "bin scan", "scatter+rank", "new generation", "cumsum+lag+outputs"]
for k, nm in enumerate(names):
    print("  %-20s %7.2f us/step" % (nm, c[k] / (nobs - 1) / 1965.0))
