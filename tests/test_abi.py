"""The C-ABI library loads on a box without a GPU and exports every symbol include/b200fit.h declares.
No compute calls here (those are the -m gpu tests)."""
import ctypes as C
import os
import re

import pytest

from mufasa_b200 import _abi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    text = open(os.path.join(ROOT, "include", "b200fit.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(b200fit_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_exported_and_bound():
    lib = _abi.load()
    names = _declared()
    assert len(names) >= 12
    for n in names:
        assert hasattr(lib, n), "libb200fit.so does not export %s" % n
        assert n in _abi.SIGNATURES, "no ctypes prototype for %s" % n
    assert lib.b200fit_abi_version() == _abi.ABI_VERSION


def test_problem_struct_and_sizes():
    lib = _abi.load()
    assert lib.b200fit_nchan_stride(800) == 800 and lib.b200fit_nchan_stride(801) == 832
    p1 = _abi.make_problem("nh3_multi_v", 1, 800, 1000, -29.0, 0.0724, [0] * 4, [1] * 4)
    p2 = _abi.make_problem("n2hp_multi_v", 2, 1200, 1000, -30.0, 0.05, [0] * 8, [1] * 8)
    w1, w2 = lib.b200fit_workspace_bytes(C.byref(p1)), lib.b200fit_workspace_bytes(C.byref(p2))
    assert 1000 * 400 < w1 < 1000 * 1200 and 1000 * 1000 < w2 < 1000 * 2000    # lists + per-pixel LM state
    p0 = _abi.make_problem("nh3_multi_v", 1, 800, 0, -29.0, 0.0724, [0] * 4, [1] * 4)
    assert lib.b200fit_workspace_bytes(C.byref(p0)) >= 256


def test_unknown_fittype():
    from mufasa_b200.exceptions import FitTypeError
    with pytest.raises(FitTypeError):
        _abi.make_problem("co_multi_v", 1, 800, 1, 0.0, 1.0, [0] * 4, [1] * 4)


def test_no_cpu_fallback():
    """Without a CUDA device the context cannot be created and the engine raises: nothing falls back to the CPU."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    lib = _abi.load()
    ctx = C.c_void_p()
    assert lib.b200fit_create(C.byref(ctx), 0) == -3          # B200FIT_E_NOGPU
    from mufasa_b200.engine import Engine
    with pytest.raises(_abi.B200FitError):
        Engine(0)
