"""CPU (-m "not gpu"): the C-ABI library loads and exports every symbol include/gibbs_b200.h declares;
argument validation works without a GPU; the package refuses to compute without CUDA (no fallback)."""
import ctypes as C
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    hdr = open(os.path.join(ROOT, "include", "gibbs_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    return sorted(set(re.findall(r"\b(gs_[a-z0-9_]+)\s*\(", hdr)))


def test_library_exports_every_declared_symbol():
    from gibbssampler_b200 import _lib
    lib = _lib.lib()
    syms = header_symbols()
    assert len(syms) >= 35
    for s in syms:
        assert hasattr(lib, s), "missing export " + s
    assert sorted(_lib.SIGNATURES) == syms, "ctypes signature table out of sync with the header"
    assert lib.gs_version() >= 100


def test_argument_validation_without_gpu():
    from gibbssampler_b200 import _lib
    lib = _lib.lib()
    p = C.c_void_p()
    assert lib.gs_plan_create(C.byref(p), 0, 8, -1) == -1          # GS_E_BADARG: nside < 1
    assert b"nside" in lib.gs_last_error_string()
    assert lib.gs_plan_create(C.byref(p), 4, 100, -1) == -1        # lmax > 4 nside
    assert lib.gs_real_to_complex(None, None, 4, None) == -1
    assert lib.gs_alm2map_spin2(None, None, None, 0, None, None, None, None) == -1
    assert b"null plan" in lib.gs_last_error_string()
    assert lib.gs_cr_pcg_pol(None, None, None, None, None, 0.0, None, None, None, None, 0, 1e-5, 10, 8, None, None, None) == -1


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from gibbssampler_b200 import GibbsB200Error
    from gibbssampler_b200.sht import Plan
    with pytest.raises(GibbsB200Error):
        Plan(4, 8)
    from gibbssampler_b200 import utils
    import numpy as np
    with pytest.raises(GibbsB200Error):
        utils.real_to_complex(np.zeros(9))


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "gibbssampler_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                for pat in (r"^\s*import\s+oracle", r"^\s*from\s+oracle", r"from\s+\.\.?\s*oracle", r"liboracle", r"oracle/_"):
                    assert not re.search(pat, src, flags=re.M), f + " uses the oracle: " + pat
