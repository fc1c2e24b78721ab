#!/usr/bin/env python
"""Per-CTA phase clocks of the exchange kernel: usage phase_clocks.py N [NOBS]."""
import sys, os, ctypes
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tools"))
import numpy as np, torch
import probe_sv
from pmmh_qn_b200 import kernels as K, _lib
K.set_sv_algorithm(2)
n = int(sys.argv[1]); nobs = int(sys.argv[2]) if len(sys.argv) > 2 else 1001
probe_sv.run(n, nobs=nobs, reps=1)
buf = torch.zeros((148, 16), dtype=torch.int64, device="cuda:0")
_lib.load().pmmh_sv_debug_profile(ctypes.c_void_p(buf.data_ptr()))
r = probe_sv.run(n, nobs=nobs, reps=0)
torch.cuda.synchronize()
_lib.load().pmmh_sv_debug_profile(None)
c = buf.cpu().numpy().astype(np.float64)
names = ["cumsum", "exch2+lag/2", "bk:lut", "ranges", "children", "exch1+lag/2", "B:pass1", "B:sort", "-", "tail", "bk:prefix", "bk:outputs", "bk:chunks", "bk:cdf", "bk:splitters"]
steps = nobs - 1
tot = c.sum(axis=1)
print("per-step microseconds at 1.9 GHz (mean / min / max over CTAs); total per CTA %.1f ms" % (tot.mean() / 1.9e6))
for k, nm in enumerate(names):
    v = c[:, k] / steps / 1900.0
    print("  %-10s %7.2f %7.2f %7.2f" % (nm, v.mean(), v.min(), v.max()))
work = c[:, [0, 2, 3, 4, 6, 7, 10, 11, 12, 13, 14]].sum(axis=1) / steps / 1900.0
print("  work sum   %7.2f %7.2f %7.2f" % (work.mean(), work.min(), work.max()))
