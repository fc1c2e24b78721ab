#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_sv_exchange.py -x -q -m gpu -k "chain" > gpurun_out/r2y_tests.log 2>&1; echo "tests exit $?" >> gpurun_out/r2y_tests.log
tail -5 gpurun_out/r2y_tests.log
timeout 300 python tools/bench_aux.py chains 2>&1 | tail -3
python - <<'PY'
import sys, os, ctypes
sys.path.insert(0, "tools"); sys.path.insert(0, ".")
import numpy as np, torch
import probe_sv
from pmmh_qn_b200 import kernels as K, _lib
K.set_sv_algorithm(3)
for hess in (False, True):
    probe_sv.run(4096, nobs=1001, reps=1, hess=hess)
    buf = torch.zeros((148, 16), dtype=torch.int64, device="cuda:0")
    _lib.load().pmmh_sv_debug_profile(ctypes.c_void_p(buf.data_ptr()))
    probe_sv.run(4096, nobs=1001, reps=0, hess=hess)
    torch.cuda.synchronize()
    _lib.load().pmmh_sv_debug_profile(None)
    c = buf.cpu().numpy().astype(np.float64)[0]
    names = ["resample+propagate", "bin scan", "scatter+rank", "new generation", "cumsum+lag+outputs"]
    for k, nm in enumerate(names):
        print("  hess %d %-20s %7.2f us/step" % (hess, nm, c[k] / 1000 / 1965.0))
PY
