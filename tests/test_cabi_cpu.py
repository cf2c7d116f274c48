"""CPU-only checks of the C-ABI library: it loads, exports every symbol include/hop_b200.h declares,
and refuses to compute without a CUDA device (no CPU fallback)."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from hop import _cabi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "hop_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(hop_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_are_exported_and_bound():
    lib = _cabi.load()
    declared = _declared_symbols()
    assert len(declared) >= 12
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in hop_b200.h but not exported"
        assert name in _cabi.SIGNATURES, f"{name} has no ctypes signature in hop/_cabi.py"
    assert sorted(_cabi.SIGNATURES) == declared
    assert lib.hop_abi_version() == 1
    assert b"sm_100a" in lib.hop_version()


def test_supported_dims_table():
    lib = _cabi.load()
    for d, m in ((3, 1), (4, 2), (5, 1), (12, 4), (13, 4)):
        assert lib.hop_select_supported(d, m) == 1
    assert lib.hop_select_supported(7, 3) == 0


def test_bad_arguments_are_rejected_before_any_device_work():
    lib = _cabi.load()
    rc = lib.hop_select_f64(1, 8, 13, 4, 5, 9, None, None, None, None, 0, None, None, None, 0, None, None, None, None, None)
    assert rc == -1 and b"T_max" in lib.hop_last_error_string()       # T_max > N


@pytest.mark.skipif(_cabi.load().hop_device_count() > 0, reason="a GPU is present")
def test_no_cpu_fallback_without_a_device():
    lib = _cabi.load()
    rc = lib.hop_select_f64(1, 8, 13, 4, 1, 8, None, None, None, None, 0, None, None, None, 0, None, None, None, None, None)
    assert rc == -3 and b"no CPU fallback" in lib.hop_last_error_string()
    from hop import api
    with pytest.raises(_cabi.HopError):
        api.select_horizon_host(__import__("hop").cases.make_case("Quadrotor", N=128), np.zeros((2, 12)))
