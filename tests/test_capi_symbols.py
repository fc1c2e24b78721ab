"""CPU: the C-ABI library builds (nvcc cross-compiles), loads, and exports every symbol that
include/pmmh_qn.h declares; the product path fails loudly without a GPU (no CPU fallback)."""
import ctypes
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "pmmh_qn.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(pmmh_[a-z0-9_]+)\s*\(", text)))


def test_header_declares_the_expected_entry_points():
    syms = _declared_symbols()
    for must in ("pmmh_flps_sv_corr", "pmmh_bpf_sv_corr", "pmmh_importance_discrete",
                 "pmmh_crank_nicolson", "pmmh_subsample_indices", "pmmh_logistic_loglike",
                 "pmmh_split_rvs", "pmmh_norm_cdf", "pmmh_sv_workspace_bytes",
                 "pmmh_flps_sv_corr_host", "pmmh_bpf_sv_corr_host",
                 "pmmh_importance_discrete_host", "pmmh_stratified_host",
                 "pmmh_svsplit_init", "pmmh_svsplit_weights", "pmmh_svsplit_children", "pmmh_svsplit_plan",
                 "pmmh_svsplit_pack", "pmmh_svsplit_pack_direct", "pmmh_svsplit_sort", "pmmh_svsplit_tail", "pmmh_svsplit_finish",
                 "pmmh_flps_sv_corr_philox", "pmmh_flps_sv_corr_streamed"):
        assert must in syms


def test_library_builds_loads_and_exports_every_declared_symbol():
    import pmmh_qn_b200
    path = pmmh_qn_b200.build()
    assert os.path.exists(path)
    lib = ctypes.CDLL(path)
    for name in _declared_symbols():
        assert hasattr(lib, name), "missing export: " + name
    from pmmh_qn_b200 import _lib
    assert set(_lib.SIGNATURES) == set(_declared_symbols())
    assert _lib.load().pmmh_version() >= 100


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from pmmh_qn_b200 import _lib, kernels as K
    from toy_models import ToySVModel
    import golden_inputs as gi
    with pytest.raises(_lib.PmmhError):
        K.norm_cdf(torch.zeros(4, dtype=torch.float64))
    with pytest.raises(_lib.PmmhError):
        K.flps_sv_corr(torch.zeros(20, dtype=torch.float64), torch.zeros(4, dtype=torch.float64),
                       torch.zeros(20, dtype=torch.float64), torch.zeros((20, 8), dtype=torch.float64))
    from pmmh_qn_b200 import ParticleMethodsCUDA
    with pytest.raises(RuntimeError):
        ParticleMethodsCUDA(ToySVModel(gi.sv_obs(20), gi.SV_PARAM_SETS[0]))
    from pmmh_qn_b200.state.particle_methods.split import SplitParticleMethodsCUDA
    with pytest.raises(RuntimeError):
        SplitParticleMethodsCUDA(ToySVModel(gi.sv_obs(40), gi.SV_PARAM_SETS[0]), 1000, local_world=2)
    with pytest.raises(_lib.PmmhError):
        K.flps_sv_corr_philox(torch.zeros(40, dtype=torch.float64), torch.zeros(4, dtype=torch.float64),
                              torch.zeros(40, dtype=torch.float64), 1, 0, 1000)
    # the C ABI itself reports the missing device instead of computing anything
    nbytes = ctypes.c_size_t()
    rc = _lib.load().pmmh_sv_workspace_bytes(361, 75, 10, 1, 0, 0, 0, 0, ctypes.byref(nbytes))
    assert rc != 0 and _lib.load().pmmh_last_error()


def test_product_package_never_imports_the_oracle():
    """The oracle is test infrastructure: no file under pmmh-qn_b200/ may reference it."""
    pkg = os.path.join(ROOT, "pmmh-qn_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in text and "pmmh_oracle" not in text.replace(
                    "oracle/pmmh_oracle.c", ""), os.path.join(dirpath, f)
