"""CPU checks of the drop-in boundary: libcvo_b200.so loads, exports every symbol include/cvo_b200.h
declares, its parameter defaults equal the two reference constructors, and it fails loudly without a GPU
(no CPU fallback).  No compute calls here."""
import ctypes
import os
import re

import pytest

from cvo_rgbd_b200 import capi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "cvo_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(cvo_b200_[a-z_0-9]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    lib = capi.load()
    declared = _declared_symbols()
    assert len(declared) >= 19
    for sym in declared:
        assert hasattr(lib, sym), "libcvo_b200.so does not export %s" % sym
    assert sorted(capi.EXPORTS) == declared


def test_params_struct_layout_matches_header():
    # 2 ints, 3 floats, (pad), double, 7 floats, int, 4 floats, int  -> the C compiler's layout
    assert ctypes.sizeof(capi.Params) == 88
    assert capi.Params.dl_step.offset == 24 and capi.Params.fixed_iters.offset == 80
    assert ctypes.sizeof(capi.IterRec) == 152


def test_default_params_equal_reference_constructors():
    p = capi.default_params("cvo")  # src/cvo.cpp:18-48
    assert (p.mode, p.ell_policy, p.max_iter, p.fixed_iters) == (capi.MODE_CVO, capi.ELL_SCHEDULE, 2000, 0)
    assert p.ell_init == pytest.approx(0.15) and p.sigma == pytest.approx(0.1) and p.sp_thres == pytest.approx(8e-3)
    assert (p.c, p.d, p.c_ell, p.c_sigma) == (7.0, 7.0, 200.0, 1.0)
    assert p.min_step == pytest.approx(0.2) and p.max_step == pytest.approx(0.8)
    assert p.eps == pytest.approx(5e-5) and p.eps_2 == pytest.approx(1e-5)
    a = capi.default_params("acvo")  # src/adaptive_cvo.cpp:18-50
    assert (a.mode, a.ell_policy) == (capi.MODE_ACVO, capi.ELL_ADAPTIVE)
    assert a.ell_init == pytest.approx(0.1) and a.ell_min == pytest.approx(0.0391) and a.ell_max == pytest.approx(0.15)
    assert a.dl_step == pytest.approx(0.3) and a.sp_thres == pytest.approx(8.315e-3)
    assert a.c_ell == pytest.approx(0.5) and a.c_sp_thres == pytest.approx(8.315e-3)


def test_defaults_agree_with_the_oracle_defaults():
    from oracle import cvo_oracle as O
    for kind in ("cvo", "acvo"):
        g, o = capi.default_params(kind), O.default_params(kind)
        for name, _ in capi.Params._fields_:
            assert getattr(g, name) == getattr(o, name), (kind, name)


def test_no_cpu_fallback_without_a_device():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(capi.CvoB200Error):
        capi.Context(0, 1024, 1)


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "cvo_rgbd_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".hpp", ".cpp")):
                src = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in src.replace("# oracle-free", ""), os.path.join(dirpath, f)
    for f in os.listdir(os.path.join(ROOT, "include")):
        assert "oracle" not in open(os.path.join(ROOT, "include", f)).read()
