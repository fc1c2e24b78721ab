"""CPU, build container only (skipped where /root/reference is absent): the reference's unmodified
quasi-Newton sampler and Cython estimator run under the lock-step front-end (parameter/lockstep.py).
The check itself runs in a subprocess: importing the reference changes global interpreter state."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.skipif(not os.path.isdir("/root/reference/python"), reason="needs the reference checkout")
def test_reference_sampler_under_the_lockstep_front_end():
    proc = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "lockstep_reference_check.py")],
                          stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=600)
    assert proc.returncode == 0, proc.stderr[-2000:]
    out = json.loads(proc.stdout.strip().splitlines()[-1])
    assert out["single_chain_equals_recorded_run"], out
    assert out["three_chains_lockstep_equal_solo"], out
    assert out["batch_sizes"] and max(out["batch_sizes"]) == 3, out
