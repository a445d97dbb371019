"""The C-ABI library loads and exports every symbol include/amg_b200.h declares (no compute calls)."""
import ctypes as C
import os
import re

from amg_b200 import capi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_functions():
    src = open(os.path.join(ROOT, "include", "amg_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    names = re.findall(r"\b((?:SSS|amgb200)_[A-Za-z0-9_]+)\s*\(", src)
    return sorted(set(n for n in names if not n.endswith("_SA")))


def test_library_exports_every_declared_symbol():
    lib = C.CDLL(capi.LIB_PATH)
    decl = declared_functions()
    assert len(decl) >= 30
    missing = [n for n in decl if not hasattr(lib, n)]
    assert not missing, f"declared in include/amg_b200.h but not exported: {missing}"


def test_python_binding_list_matches_header():
    assert sorted(capi.EXPORTS) == declared_functions()


def test_struct_layouts_match_reference_lp64():
    # sizes/offsets measured against amg/SSS_main.h with gcc
    assert C.sizeof(capi.Mat) == 40 and capi.Mat.row_ptr.offset == 16 and capi.Mat.val.offset == 32
    assert C.sizeof(capi.Vec) == 16 and C.sizeof(capi.IVec) == 16 and C.sizeof(capi.Rtn) == 24
    assert C.sizeof(capi.Pars) == 104 and capi.Pars.tol.offset == 8 and capi.Pars.trunc_threshold.offset == 96
    assert C.sizeof(capi.Comp) == 184 and capi.Comp.R.offset == 40 and capi.Comp.P.offset == 80
    assert capi.Comp.b.offset == 120 and capi.Comp.x.offset == 136 and capi.Comp.cfmark.offset == 152 and capi.Comp.wp.offset == 168
    assert C.sizeof(capi.Amg) == 144 and capi.Amg.cg.offset == 8 and capi.Amg.pars.offset == 16 and capi.Amg.rtn.offset == 120
    assert C.sizeof(capi.Smtr) == 72 and capi.Smtr.ordering.offset == 64


def test_default_parameters_match_reference_main():
    p = capi.Pars()
    capi.lib().amgb200_default_pars(C.byref(p))      # amg/SSS_main.c:25-64
    assert (p.smoother, p.max_it, p.max_levels, p.coarse_dof, p.cycle_type, p.cf_order) == (2, 100, 30, 10, 1, 1)
    assert (p.pre_iter, p.post_iter, p.cs_type, p.interp_type) == (2, 2, 1, 1)
    assert (p.tol, p.ctol, p.max_row_sum, p.strong_threshold, p.trunc_threshold) == (1e-6, 1e-7, 0.9, 0.3, 0.2)


def test_version_string():
    assert b"sm_100a" in capi.lib().amgb200_version()
