"""The C-ABI library loads on a CPU-only box and exports exactly what include/dsoft.h declares.
No compute entry point is called here (that needs a GPU); argument validation paths are."""
import ctypes as C
import os
import re

import pytest

from conftest import ROOT


def header_functions():
    text = open(os.path.join(ROOT, "include", "dsoft.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(dsoft_[a-z0-9_]+)\s*\(", text)))


def test_header_declares_expected_surface():
    names = header_functions()
    for must in ("dsoft_plan_create", "dsoft_pack", "dsoft_forward", "dsoft_backward", "dsoft_last_error"):
        assert must in names


def test_library_exports_every_declared_symbol(pkg):
    from dinosoft_b200 import _build, _cabi

    assert os.path.exists(_build.LIB_PATH), "libdsoft.so not built (run __graft_entry__.build())"
    raw = C.CDLL(_build.LIB_PATH)
    for name in header_functions():
        assert hasattr(raw, name), f"{name} declared in include/dsoft.h but not exported"
    # and the ctypes prototype table covers the same set (no stale / missing bindings)
    assert sorted(_cabi.PROTOTYPES) == header_functions()


def test_no_device_is_reported_not_hidden(pkg):
    """On a box without an sm_100 GPU plan creation must fail loudly (no CPU fallback)."""
    import torch

    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from dinosoft_b200 import _cabi

    lib = _cabi.lib()
    assert lib.dsoft_version() >= 100
    shape = _cabi.Shape(b=128, world=1, rank=0, D=64, Dp=0, Dd=0, flags=0, teacher_temp=0.0, text_temp=0.0)
    h = C.c_void_p()
    rc = lib.dsoft_plan_create(C.byref(shape), C.byref(h))
    assert rc != 0
    assert lib.dsoft_last_error()


@pytest.mark.parametrize("bad", [dict(b=0), dict(D=12), dict(world=2, rank=2), dict(flags=2), dict(flags=1, Dd=0)])
def test_plan_argument_validation(pkg, bad):
    from dinosoft_b200 import _cabi

    lib = _cabi.lib()
    kw = dict(b=128, world=1, rank=0, D=64, Dp=0, Dd=64, flags=0, teacher_temp=0.15, text_temp=0.02)
    kw.update(bad)
    h = C.c_void_p()
    rc = lib.dsoft_plan_create(C.byref(_cabi.Shape(**kw)), C.byref(h))
    assert rc == -1, (bad, rc, lib.dsoft_last_error())
