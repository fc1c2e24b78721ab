#!/usr/bin/env python
"""BASELINE config 2 sweep: SV smoother T=1000, N = 2^10 .. 2^20, gradient and gradient + Hessian,
u resident in HBM.  One JSON line per case."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tools"))
import probe_sv
for n in (1 << 10, 1 << 12, 1 << 14, 1 << 16, 1 << 18, 1 << 20):
    probe_sv.run(n, hess=False, reps=3)
    probe_sv.run(n, hess=True, reps=2)
