#!/usr/bin/env python
"""Development probe: time the SV smoother kernels at several N and decode the diagnostics.
usage: probe_fast.py [algo] N [N ...]   (algo: 0 auto, 1 general, 2 exchange without fallback)"""
import sys, os, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tools"))
import probe_sv
from pmmh_qn_b200 import kernels as K
algo = int(sys.argv[1])
K.set_sv_algorithm(algo)
for n in [int(x) for x in sys.argv[2:]]:
    r = probe_sv.run(n, reps=3)
    d = r["diag"]; info = d[7]
    print("   algo=%d kernel=%d status=%d reason=%d step=%d maxchunk=%d" % (algo, d[6], d[2], info & 255, (info >> 8) & 0xffffff, d[1]), flush=True)
