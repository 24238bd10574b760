"""The C-ABI library loads on a CPU-only box and exports every symbol include/*.h declares."""
import ctypes
import os
import re

import pytest

from conftest import ROOT
from rmt_app_b200 import capi


def _declared_symbols():
    names = set()
    for fn in os.listdir(os.path.join(ROOT, "include")):
        if not fn.endswith(".h"):
            continue
        src = open(os.path.join(ROOT, "include", fn)).read()
        src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
        names |= set(re.findall(r"\b(rmt_[a-z0-9_]+)\s*\(", src))
    return names


def test_every_declared_symbol_is_exported_and_bound():
    lib = capi.lib()
    decl = _declared_symbols()
    assert len(decl) >= 20
    for name in decl:
        assert hasattr(lib, name), "librmtb200.so does not export %s" % name
    assert decl == set(capi.SIGNATURES), decl ^ set(capi.SIGNATURES)


def test_version_and_no_cpu_fallback():
    lib = capi.lib()
    assert b"sm_100a" in lib.rmt_version()
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(capi.RmtError, match="no CPU fallback|CUDA"):
        capi.init(0)
    from rmt_app_b200 import rmtExe
    import cases
    with pytest.raises(capi.RmtError, match="no CPU fallback"):
        rmtExe(cases.methanol_readme_input())


def test_compute_entry_points_fail_loudly_without_context():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    lib = capi.lib()
    assert lib.rmt_n1_rhs(ctypes.c_uint64(1), 4, None, None, None, None) != 0
    assert lib.rmt_last_error()
