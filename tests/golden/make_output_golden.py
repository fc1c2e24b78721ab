#!/usr/bin/env python
"""tests/golden/output_writer.json: the files the REFERENCE's own writer (parameter/mcmc/output.py:266-356,
helpers/file_system.py:43-77) produces for the seeded stand-in sampler of output_inputs.py.  Build
container only (imports /root/reference with the shims of make_golden.py)."""
import contextlib
import gzip
import io
import json
import os
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import make_golden  # noqa: E402
import output_inputs  # noqa: E402

make_golden.import_reference_python()
from parameter.mcmc import output as ref_output  # noqa: E402

golden = {}
for tag, bench in (("plain", False), ("benchmark", True)):
    smp = output_inputs.fake_sampler(benchmark=bench)
    with tempfile.TemporaryDirectory() as tmp, contextlib.redirect_stdout(io.StringIO()):
        ref_output.save_to_file(smp, tmp, sim_name="sim", sim_desc="a description")
        files = {}
        for fn in sorted(os.listdir(os.path.join(tmp, "sim"))):
            with gzip.open(os.path.join(tmp, "sim", fn), "rt") as fh:
                d = json.loads(fh.read())
            for k in ("simulation_time", "time"):
                if k in d:
                    d[k] = "<time>"
            files[fn] = d
    golden[tag] = files
with open(os.path.join(HERE, "output_writer.json"), "w") as fh:
    json.dump(golden, fh, indent=0, sort_keys=True)
print({k: sorted(v) for k, v in golden.items()})
