#!/usr/bin/env python
"""BASELINE config 2 sweep: SV smoother T=1000, N = 2^10 .. 2^20, gradient and gradient + Hessian,
u resident in HBM, automatic kernel selection; plus the grid kernel forced (algorithm 6) and the exchange
kernel forced (algorithm 2) at the sizes where the automatic threshold sits.  One JSON line per case."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tools"))
import probe_sv
from pmmh_qn_b200 import kernels as K
for n in (1 << 10, 1 << 12, 1 << 13, 1 << 14, 1 << 15, 1 << 16, 1 << 17, 1 << 18, 1 << 19, 1 << 20):
    probe_sv.run(n, hess=False, reps=3)
    if n in (1 << 10, 1 << 12, 1 << 14, 1 << 16, 1 << 18, 1 << 20):
        probe_sv.run(n, hess=True, reps=2)
for alg in (6, 2):
    K.set_sv_algorithm(alg)
    for n in (1 << 13, 1 << 14, 1 << 15, 1 << 16, 1 << 17):
        try:
            r = probe_sv.run(n, hess=False, reps=3)
        except Exception as e:
            print('{"alg": %d, "n": %d, "error": "%s"}' % (alg, n, str(e)[:80]))
# the Hessian branch around its automatic threshold (2^16): grid kernel forced (6) against the general kernel (1)
for alg in (6, 1):
    K.set_sv_algorithm(alg)
    for n in (1 << 14, 1 << 15, 1 << 16, 1 << 17):
        try:
            r = probe_sv.run(n, hess=True, reps=2)
        except Exception as e:
            print('{"alg": %d, "n": %d, "hess": 1, "error": "%s"}' % (alg, n, str(e)[:80]))
K.set_sv_algorithm(0)
