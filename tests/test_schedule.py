"""CPU validation of the device-side analysis (amg_b200/csrc/analysis.cpp): the wavefront schedule of
the ordered Gauss-Seidel and the SELL / CSR device layouts.  The debug walk executes the rows
wavefront by wavefront in REVERSE intra-wavefront order through the device layout; if the schedule is
right the result is bit-identical to the reference's sequential sweep."""
import ctypes as C

import numpy as np
import pytest

import oracle_ffi
from amg_b200 import HostHierarchy, capi, generate

L = oracle_ffi.test_hooks()
L.amgb200_debug_gs_walk.argtypes = [C.POINTER(capi.Mat), capi.c_int_p, C.c_int, C.c_int, capi.c_double_p, capi.c_double_p]
L.amgb200_debug_schedule.restype = C.c_int
L.amgb200_debug_schedule.argtypes = [C.POINTER(capi.Mat), capi.c_int_p, capi.c_int_p, capi.c_int_p, C.c_int, capi.c_int_p]
L.amgb200_debug_spmv_walk.argtypes = [C.POINTER(capi.Mat), C.c_int, capi.c_double_p, capi.c_double_p]

CASES = [("p2d", 48, 0.0), ("p3d", 16, 0.0), ("aniso3d", 24, 1e-3), ("v27", 10, 0.0)]


def schedule(mat, mark):
    n = mat.num_rows
    order = np.zeros(n, np.int32); wf = np.zeros(n + 2, np.int32); cnt = np.zeros(4, np.int32)
    W = L.amgb200_debug_schedule(C.byref(mat), capi.iptr(mark), capi.iptr(order), capi.iptr(wf), n + 2, capi.iptr(cnt))
    return order, wf[:W + 1], cnt


@pytest.mark.parametrize("case", CASES)
def test_wavefront_schedule_properties(case):
    hier = HostHierarchy(generate(*case), tol=1e-8)
    for l in range(hier.num_levels - 1):
        c = hier.level(l)
        n = c.A.num_rows
        mark = np.ascontiguousarray(hier.cfmark(l))
        order, wf, cnt = schedule(c.A, mark)
        assert sorted(order.tolist()) == list(range(n)), "schedule is not a permutation"
        assert wf[0] == 0 and wf[-1] == n and (np.diff(wf) > 0).all()
        nF = int((mark != 1).sum())
        assert wf[cnt[0]] == nF, "F pass must come first and hold every non-C row"
        assert (mark[order[:nF]] != 1).all() and (mark[order[nF:]] == 1).all()
        # rows inside a wavefront are in ascending natural order and mutually uncoupled
        rp, ci, _ = capi.mat_arrays(c.A)
        pos = np.empty(n, np.int64); pos[order] = np.arange(n)
        wf_of = np.searchsorted(wf, pos, side="right") - 1
        for i in range(n):
            for j in ci[rp[i]:rp[i + 1]]:
                if j != i and (mark[i] == 1) == (mark[j] == 1):
                    assert wf_of[i] != wf_of[j], "coupled same-pass rows share a wavefront"
                    assert (wf_of[j] < wf_of[i]) == (j < i), "wavefront order must follow row order along every coupling"
        assert cnt[3] == 0


@pytest.mark.parametrize("case", CASES)
@pytest.mark.parametrize("kind", [0, 1, 2, 3])   # SELL-32 | CSR | packed per-wavefront blocks of the streaming smoother (two-phase rows) | per-(wavefront, CTA) blocks of the cluster streaming smoother
def test_wavefront_execution_is_bit_identical_to_sequential_sweep(case, kind, oracle):
    hier = HostHierarchy(generate(*case), tol=1e-8)
    for l in range(hier.num_levels - 1):
        c = hier.level(l)
        n = c.A.num_rows
        mark = np.ascontiguousarray(hier.cfmark(l))
        rng = np.random.default_rng(l)
        x0, b = rng.standard_normal(n), rng.standard_normal(n)
        want = oracle.gs_cf(c.A, mark, x0, b, 2, 1)
        x = x0.copy()
        L.amgb200_debug_gs_walk(C.byref(c.A), capi.iptr(mark), kind, 2, capi.dptr(x), capi.dptr(b))
        assert x.tobytes() == want.tobytes()


def test_poisson_level0_is_two_colour():
    for kind, N in (("p2d", 32), ("p3d", 12)):
        hier = HostHierarchy(generate(kind, N), tol=1e-8)
        _, _, cnt = schedule(hier.level(0).A, np.ascontiguousarray(hier.cfmark(0)))
        assert (cnt[0], cnt[1]) == (1, 1)          # exact red/black split: each pass is one wavefront


def test_nonsymmetric_pattern_keeps_read_old_value_order(oracle):
    """a_ij stored but a_ji not: row j > i must still run after row i has read the old x_j"""
    rp = np.array([0, 3, 4, 5, 6], np.int32)
    ci = np.array([0, 1, 2, 1, 2, 3], np.int32)
    va = np.array([4.0, -1.0, -1.0, 3.0, 5.0, 2.0])
    mat, keep = capi.mat_from_arrays(rp, ci, va, 4)
    mark = np.zeros(4, np.int32)
    order, wf, cnt = schedule(mat, mark)
    assert cnt[2] == 0                                         # pattern reported as non-symmetric
    x0 = np.array([1.0, 2.0, 3.0, 4.0]); b = np.array([1.0, 1.0, 1.0, 1.0])
    want = oracle.gs_cf(mat, mark, x0, b, 3, 1)
    for kind in (0, 1, 2, 3):
        x = x0.copy()
        L.amgb200_debug_gs_walk(C.byref(mat), capi.iptr(mark), kind, 3, capi.dptr(x), capi.dptr(b))
        assert x.tobytes() == want.tobytes()


@pytest.mark.parametrize("kind", [0, 1])
def test_layout_spmv_walk(kind, oracle):
    hier = HostHierarchy(generate("v27", 8), tol=1e-8)
    for l in range(hier.num_levels):
        for which in ("A", "P", "R"):
            if which != "A" and l == hier.num_levels - 1:
                continue
            m = getattr(hier.level(l), which)
            x = np.random.default_rng(3).standard_normal(m.num_cols)
            y = np.zeros(m.num_rows)
            L.amgb200_debug_spmv_walk(C.byref(m), kind, capi.dptr(x), capi.dptr(y))
            assert y.tobytes() == oracle.mxy(m, x).tobytes()


L.amgb200_debug_fused_walk.restype = C.c_int
L.amgb200_debug_fused_walk.argtypes = [C.POINTER(capi.Mat), capi.c_int_p, C.POINTER(capi.Mat), capi.c_int_p, C.c_int, C.POINTER(capi.Mat),
                                       C.c_int, C.c_int, C.c_int, capi.c_double_p, capi.c_double_p, capi.c_double_p, capi.c_double_p]


@pytest.mark.parametrize("case", CASES)
@pytest.mark.parametrize("nch,lag,tickets", [(1, 0, 32), (5, 1, 32), (16, 0, 3), (64, 1, 32), (64, 3, 8), (200, 1, 1)])
def test_fused_residual_restriction_ticket_list(case, nch, lag, tickets, oracle):
    """the ticket list of the fused residual (+) restriction launch (amg/Solve/SSS_cycle.cu:916-921), executed on the CPU through the
    device layouts: every ticket depends on smaller tickets only, every slice runs once, the declared chunk ranges cover what a
    slice of R reads, and r / b_{l+1} are the reference's bit for bit"""
    hier = HostHierarchy(generate(*case), tol=1e-8)
    for l in range(hier.num_levels - 1):
        c, cc = hier.level(l), hier.level(l + 1)
        n, ncoarse = c.A.num_rows, cc.A.num_rows
        coarsest = l + 1 == hier.num_levels - 1
        mark = np.ascontiguousarray(hier.cfmark(l))
        markc = None if coarsest else np.ascontiguousarray(hier.cfmark(l + 1))
        rng = np.random.default_rng(40 + l)
        x, b = rng.standard_normal(n), rng.standard_normal(n)
        r, bc = np.zeros(n), np.zeros(ncoarse)
        rc = L.amgb200_debug_fused_walk(C.byref(c.A), capi.iptr(mark), C.byref(cc.A), capi.iptr(markc) if markc is not None else None, int(coarsest),
                                        C.byref(c.R), nch, lag, tickets, capi.dptr(x), capi.dptr(b), capi.dptr(r), capi.dptr(bc))
        assert rc == 0, f"level {l}: invariant {rc} of the ticket list violated"
        want_r = oracle.amxpy(-1.0, c.A, x, b)
        assert r.tobytes() == want_r.tobytes()
        assert bc.tobytes() == oracle.mxy(c.R, want_r).tobytes()
