#!/usr/bin/env python
"""Build the UNMODIFIED reference Cython kernels into ``oracle/_ref/``.

TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` is on the product path.

The reference (compops/pmmh-qn) fixes its problem sizes at compile time with
``DEF NPART / DEF NOBS / DEF LAG`` lines:

* python/state/particle_methods/stochastic_volatility.pyx:17-19
* python/state/importance_sampling/random_effects.pyx:9-10
* python/state/direct/subsampling.pyx:5-6

This script reads the ``.pyx`` files where they lie under ``/root/reference``,
substitutes ONLY those ``DEF`` lines in a scratch copy under ``oracle/_ref/_src``,
runs Cython + gcc with the flags the reference's own ``python/setup.py`` would
use (distutils defaults: ``-O2``, no ``-ffast-math``, no ``-march``), writes one
extension module per size variant into ``oracle/_ref/`` and removes the scratch
sources again.  No reference source is ever copied into the tracked tree
(``oracle/_ref/`` is git-ignored; the built ``.so`` files travel to the GPU box).

Cython 3 needs ``legacy_implicit_noexcept=True`` to accept the reference's
``qsort`` comparator (stochastic_volatility.pyx:32,43).

Usage:  python oracle/build_ref.py            # all default variants
        python oracle/build_ref.py sv:75:361:10
"""
import os
import re
import shutil
import subprocess
import sys
import sysconfig

HERE = os.path.dirname(os.path.abspath(__file__))
REF_ROOT = os.environ.get("PMMH_REFERENCE_ROOT", "/root/reference")
OUT = os.path.join(HERE, "_ref")

SOURCES = {
    "sv": "python/state/particle_methods/stochastic_volatility.pyx",
    "re": "python/state/importance_sampling/random_effects.pyx",
    "ss": "python/state/direct/subsampling.pyx",
}

# (kind, NPART, NOBS, LAG)   -- LAG only meaningful for "sv"
DEFAULT_VARIANTS = [
    ("sv", 75, 361, 10),     # shipped constants
    ("sv", 37, 50, 10),      # N < NOBS, tiny
    ("sv", 200, 120, 10),    # N > NOBS (exercises Q7/Q10 slot>0 reads)
    ("sv", 64, 40, 4),       # different LAG
    ("sv", 1024, 1001, 10),  # BASELINE config 2, smallest N
    ("sv", 4096, 1001, 10),  # BASELINE config 4 per-chain size
    ("re", 100, 100, 0),     # shipped constants
    ("re", 64, 37, 0),       # NPART != NOBS
    ("ss", 5500, 110000, 0),  # shipped constants
    ("ss", 77, 1000, 0),
]


def module_name(kind, npart, nobs, lag):
    if kind == "sv":
        return "ref_sv_n%d_t%d_l%d" % (npart, nobs, lag)
    return "ref_%s_n%d_t%d" % (kind, npart, nobs)


def build_variant(kind, npart, nobs, lag, force=False):
    name = module_name(kind, npart, nobs, lag)
    suffix = sysconfig.get_config_var("EXT_SUFFIX")
    target = os.path.join(OUT, name + suffix)
    if os.path.exists(target) and not force:
        return target
    src_path = os.path.join(REF_ROOT, SOURCES[kind])
    if not os.path.exists(src_path):
        raise FileNotFoundError("reference source not present: " + src_path)
    with open(src_path) as fh:
        text = fh.read()
    text, n1 = re.subn(r"(?m)^DEF NPART = \d+\s*$", "DEF NPART = %d" % npart, text)
    text, n2 = re.subn(r"(?m)^DEF NOBS = \d+\s*$", "DEF NOBS = %d" % nobs, text)
    assert n1 == 1 and n2 == 1, "DEF lines not found in " + src_path
    if kind == "sv":
        text, n3 = re.subn(r"(?m)^DEF LAG = \d+\s*$", "DEF LAG = %d" % lag, text)
        assert n3 == 1
    scratch = os.path.join(OUT, "_src")
    os.makedirs(scratch, exist_ok=True)
    pyx = os.path.join(scratch, name + ".pyx")
    cfile = os.path.join(scratch, name + ".c")
    with open(pyx, "w") as fh:
        fh.write(text)
    try:
        subprocess.check_call(
            [sys.executable, "-m", "cython", "-3", "-X", "legacy_implicit_noexcept=True",
             pyx, "-o", cfile],
            stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
        import numpy
        cflags = (sysconfig.get_config_var("CFLAGS") or "-O2").split()
        # the reference's setup.py uses plain distutils flags; keep them, silence warnings
        cflags = [f for f in cflags if f not in ("-Wall", "-Wsign-compare", "-g")]
        cmd = (["gcc", "-shared", "-fPIC", "-w"] + cflags +
               ["-I" + sysconfig.get_paths()["include"], "-I" + numpy.get_include(),
                "-DNPY_NO_DEPRECATED_API=NPY_1_7_API_VERSION",
                cfile, "-o", target, "-lm"])
        subprocess.check_call(cmd)
    finally:
        if not os.environ.get("PMMH_KEEP_REF_SRC"):
            shutil.rmtree(scratch, ignore_errors=True)
    return target


def build_all(variants=None, force=False, quiet=False):
    """Build every variant; returns list of built paths.  Silently does nothing
    when the reference tree is absent (e.g. on the GPU box, which only uses the
    prebuilt files that travelled with the snapshot)."""
    if not os.path.isdir(REF_ROOT):
        return []
    os.makedirs(OUT, exist_ok=True)
    built = []
    for v in (variants or DEFAULT_VARIANTS):
        p = build_variant(*v, force=force)
        built.append(p)
        if not quiet:
            print("built", os.path.relpath(p, HERE))
    return built


def load(kind, npart, nobs, lag=10):
    """Import a built reference module (None if it is not there)."""
    import importlib.util
    name = module_name(kind, npart, nobs, lag if kind == "sv" else 0)
    path = os.path.join(OUT, name + sysconfig.get_config_var("EXT_SUFFIX"))
    if not os.path.exists(path):
        return None
    spec = importlib.util.spec_from_file_location(name, path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


if __name__ == "__main__":
    args = sys.argv[1:]
    force = "--force" in args
    args = [a for a in args if a != "--force"]
    if args:
        variants = []
        for a in args:
            parts = a.split(":")
            kind = parts[0]
            npart, nobs = int(parts[1]), int(parts[2])
            lag = int(parts[3]) if len(parts) > 3 else (10 if kind == "sv" else 0)
            variants.append((kind, npart, nobs, lag))
    else:
        variants = None
    build_all(variants, force=force)
