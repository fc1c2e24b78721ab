"""CPU: the result writer (pmmh-qn_b200/parameter/output.py) produces, file for file and key for key, what
the reference's own writer (parameter/mcmc/output.py:266-356, helpers/file_system.py:43-77) produced for
the same seeded sampler stand-in (tests/golden/output_writer.json, made by make_output_golden.py)."""
import gzip
import importlib
import json
import os

import numpy as np
import pytest

import output_inputs

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _same(a, b, path=""):
    if isinstance(b, dict):
        assert isinstance(a, dict) and sorted(a) == sorted(b), path
        for k in b:
            _same(a[k], b[k], path + "/" + k)
    elif isinstance(b, list):
        assert np.array_equal(np.asarray(a, dtype=object).shape, np.asarray(b, dtype=object).shape), path
        assert np.array_equal(np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64), equal_nan=True), path
    else:
        assert a == b or (isinstance(a, float) and isinstance(b, float) and np.isnan(a) and np.isnan(b)), (path, a, b)


@pytest.mark.parametrize("tag,bench", [("plain", False), ("benchmark", True)])
def test_writer_matches_the_reference_files(tmp_path, tag, bench):
    out = importlib.import_module("pmmh_qn_b200.parameter.output")
    with open(os.path.join(ROOT, "tests", "golden", "output_writer.json")) as fh:
        golden = json.load(fh)[tag]
    smp = output_inputs.fake_sampler(benchmark=bench)
    written = out.save_to_file(smp, str(tmp_path), sim_name="sim", sim_desc="a description", now="<time>")
    names = sorted(os.path.basename(p) for p in written)
    assert names == sorted(golden)
    for p in written:
        with gzip.open(p, "rt") as fh:
            got = json.loads(fh.read())
        _same(got, golden[os.path.basename(p)], os.path.basename(p))
