#!/usr/bin/env python
"""Small fixed workload for ncu captures: usage ncu_target.py <algo> <N> <NOBS> [reps]."""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tools"))
import probe_sv
from pmmh_qn_b200 import kernels as K
K.set_sv_algorithm(int(sys.argv[1]))
probe_sv.run(int(sys.argv[2]), nobs=int(sys.argv[3]), reps=int(sys.argv[4]) if len(sys.argv) > 4 else 1)
