"""ctypes access to the CHECKERS under oracle/ (test infrastructure only).

  Oracle()      oracle/liboracle.so      -- the C restatement (oracle/amg_oracle.c)
  Reference()   oracle/_ref/libsss_ref_{fix,asc}.so -- the reference's own sources compiled by
                oracle/build_ref.py; present in the build container and (as a prebuilt,
                git-ignored binary) on the GPU box, absent from a fresh clone.
"""
import contextlib
import ctypes as C
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from amg_b200 import capi  # noqa: E402

ORACLE_SO = os.path.join(ROOT, "oracle", "liboracle.so")
HOOKS_SO = os.path.join(ROOT, "tests", "libamgb200_testhooks.so")


def test_hooks():
    """tests/libamgb200_testhooks.so: CPU emulations of the wavefront schedule and the device layouts (amg_b200/csrc/debug_host.cpp),
    built next to the product but not part of it"""
    if not os.path.exists(HOOKS_SO):
        raise RuntimeError("tests/libamgb200_testhooks.so missing: run `make`")
    return C.CDLL(HOOKS_SO)
REF_DIR = os.path.join(ROOT, "oracle", "_ref")


@contextlib.contextmanager
def quiet():
    """silence C-level stdout (the reference prints tables and debug lines)"""
    sys.stdout.flush()
    saved = os.dup(1)
    devnull = os.open(os.devnull, os.O_WRONLY)
    os.dup2(devnull, 1)
    try:
        yield
    finally:
        os.dup2(saved, 1)
        os.close(saved)
        os.close(devnull)


MATGEN_SO = os.path.join(ROOT, "oracle", "libmatgen.so")


class MatGen:
    """oracle/libmatgen.so: the synthetic operators of SURVEY.md Appendix B restated in plain C (oracle/matgen.c), so that the
    reference arm of bench.py builds its input without loading libamgb200.so.  Returns (row_ptr, col_idx, val) numpy arrays."""
    KINDS = {"p2d": 0, "p3d": 1, "aniso3d": 2, "v27": 3}

    def __init__(self):
        if not os.path.exists(MATGEN_SO):
            raise RuntimeError("oracle/libmatgen.so missing: run `make`")
        self.L = C.CDLL(MATGEN_SO)

    def generate(self, kind, N, eps_z=1e-3):
        class M(C.Structure):
            _fields_ = [("num_rows", C.c_int), ("num_cols", C.c_int), ("num_nnzs", C.c_int), ("row_ptr", C.POINTER(C.c_int)),
                        ("col_idx", C.POINTER(C.c_int)), ("val", C.POINTER(C.c_double))]
        m = M()
        rc = self.L.orc_generate(self.KINDS[kind], int(N), C.c_double(float(eps_z)), C.byref(m))
        if rc != 0:
            raise ValueError(f"orc_generate({kind},{N}) failed: {rc}")
        rp = np.ctypeslib.as_array(m.row_ptr, shape=(m.num_rows + 1,)).copy()
        ci = np.ctypeslib.as_array(m.col_idx, shape=(m.num_nnzs,)).copy()
        va = np.ctypeslib.as_array(m.val, shape=(m.num_nnzs,)).copy()
        self.L.orc_mat_free(C.byref(m))
        return rp, ci, va


def have_ref():
    return os.path.exists(os.path.join(REF_DIR, "libsss_ref_fix.so"))


class Oracle:
    def __init__(self):
        if not os.path.exists(ORACLE_SO):
            raise RuntimeError("oracle/liboracle.so missing: run `make oracle/liboracle.so`")
        L = C.CDLL(ORACLE_SO)
        dp, ip = capi.c_double_p, capi.c_int_p
        L.orc_mv_mxy.argtypes = [C.POINTER(capi.Mat), dp, dp]
        L.orc_mv_amxpy.argtypes = [C.c_double, C.POINTER(capi.Mat), dp, dp]
        L.orc_gs_cf.argtypes = [dp, C.POINTER(capi.Mat), dp, C.c_int, ip, C.c_int]
        L.orc_gs.argtypes = [dp, C.c_int, C.c_int, C.c_int, C.POINTER(capi.Mat), dp, C.c_int]
        L.orc_cg.restype = C.c_int
        L.orc_cg.argtypes = [C.POINTER(capi.Mat), dp, dp, C.c_double, C.c_int, C.c_int]
        L.orc_gmres.restype = C.c_int
        L.orc_gmres.argtypes = [C.POINTER(capi.Mat), dp, dp, C.c_double, C.c_int, C.c_int]
        L.orc_coarse_solve.restype = C.c_int
        L.orc_coarse_solve.argtypes = [C.POINTER(capi.Mat), dp, dp, C.c_double, C.c_int, ip]
        L.orc_cycle.argtypes = [C.POINTER(capi.Amg), C.c_int]
        L.orc_solve.restype = capi.Rtn
        L.orc_solve.argtypes = [C.POINTER(capi.Amg), C.POINTER(capi.Vec), C.POINTER(capi.Vec), C.c_int, C.c_int, dp, C.c_int]
        L.orc_norm2.restype = C.c_double
        L.orc_norm2.argtypes = [C.c_int, dp]
        self.L = L

    def mxy(self, mat, x):
        y = np.zeros(mat.num_rows)
        xx = np.ascontiguousarray(x, np.float64)
        self.L.orc_mv_mxy(C.byref(mat), capi.dptr(xx), capi.dptr(y))
        return y

    def amxpy(self, alpha, mat, x, y):
        out = np.array(y, np.float64, copy=True)
        xx = np.ascontiguousarray(x, np.float64)
        self.L.orc_mv_amxpy(alpha, C.byref(mat), capi.dptr(xx), capi.dptr(out))
        return out

    def gs_cf(self, mat, mark, x0, b, sweeps, order=1):
        x = np.array(x0, np.float64, copy=True)
        bb = np.ascontiguousarray(b, np.float64)
        mk = np.ascontiguousarray(mark, np.int32)
        self.L.orc_gs_cf(capi.dptr(x), C.byref(mat), capi.dptr(bb), sweeps, capi.iptr(mk), order)
        return x

    def gs_natural(self, mat, x0, b, sweeps, backward=False):
        """SSS_smooth.c:90-137 as the cycle calls it: forward 0..n-1 (pre) or backward n-1..0 (post)"""
        x = np.array(x0, np.float64, copy=True)
        bb = np.ascontiguousarray(b, np.float64)
        n = mat.num_rows
        if backward:
            self.L.orc_gs(capi.dptr(x), n - 1, 0, -1, C.byref(mat), capi.dptr(bb), sweeps)
        else:
            self.L.orc_gs(capi.dptr(x), 0, n - 1, 1, C.byref(mat), capi.dptr(bb), sweeps)
        return x

    def coarse_solve(self, mat, x0, b, tol, beta_mode=0):
        x = np.array(x0, np.float64, copy=True)
        bb = np.ascontiguousarray(b, np.float64)
        its = (C.c_int * 2)()
        with quiet():
            st = self.L.orc_coarse_solve(C.byref(mat), capi.dptr(bb), capi.dptr(x), tol, beta_mode, its)
        return st, x, (its[0], its[1])

    def cycle(self, hier, x0, b, beta_mode=0):
        x = np.array(x0, np.float64, copy=True)
        bb = np.ascontiguousarray(b, np.float64)
        hier.mg.cg[0].x = capi.vec_from_array(x)
        hier.mg.cg[0].b = capi.vec_from_array(bb)
        with quiet():
            self.L.orc_cycle(C.byref(hier.mg), beta_mode)
        hier.mg.cg[0].x = capi.Vec(0, None)
        hier.mg.cg[0].b = capi.Vec(0, None)
        return x

    def solve(self, hier, x0, b, beta_mode=0, hist_cap=200):
        x = np.array(x0, np.float64, copy=True)
        bb = np.ascontiguousarray(b, np.float64)
        vx, vb = capi.vec_from_array(x), capi.vec_from_array(bb)
        hist = np.zeros(hist_cap)
        rtn = self.L.orc_solve(C.byref(hier.mg), C.byref(vx), C.byref(vb), beta_mode, 0, capi.dptr(hist), hist_cap)
        hier.mg.cg[0].x = capi.Vec(0, None)
        hier.mg.cg[0].b = capi.Vec(0, None)
        return rtn, x, hist[:rtn.nits].copy()


class Reference:
    """the reference's own objects (oracle/_ref)"""

    def __init__(self, mode="fix"):
        path = os.path.join(REF_DIR, f"libsss_ref_{mode}.so")
        if not os.path.exists(path):
            raise RuntimeError(f"{path} missing (built by oracle/build_ref.py where /root/reference exists)")
        L = C.CDLL(path)
        L.SSS_amg_setup.argtypes = [C.POINTER(capi.Amg), C.POINTER(capi.Mat), C.POINTER(capi.Pars)]
        L.SSS_amg_solve.restype = capi.Rtn
        L.SSS_amg_solve.argtypes = [C.POINTER(capi.Amg), C.POINTER(capi.Vec), C.POINTER(capi.Vec)]
        L.SSS_amg_cycle.argtypes = [C.POINTER(capi.Amg)]
        L.SSS_amg_data_destroy.argtypes = [C.POINTER(capi.Amg)]
        L.SSS_blas_mv_mxy.argtypes = [C.POINTER(capi.Mat), C.POINTER(capi.Vec), C.POINTER(capi.Vec)]
        L.SSS_blas_mv_amxpy.argtypes = [C.c_double, C.POINTER(capi.Mat), C.POINTER(capi.Vec), C.POINTER(capi.Vec)]
        L.SSS_amg_smoother_pre.argtypes = [C.POINTER(capi.Smtr)]
        L.SSS_amg_coarest_solve.argtypes = [C.POINTER(capi.Mat), C.POINTER(capi.Vec), C.POINTER(capi.Vec), C.c_double]
        self.L = L

    def setup(self, A, pars):
        mg = capi.Amg()
        p = capi.Pars.from_buffer_copy(pars)
        with quiet():
            self.L.SSS_amg_setup(C.byref(mg), C.byref(A.c), C.byref(p))
        return mg

    def destroy(self, mg):
        mg.cg[0].x = capi.Vec(0, None)
        mg.cg[0].b = capi.Vec(0, None)
        self.L.SSS_amg_data_destroy(C.byref(mg))

    def solve_history(self, mg, x0, b, tol, max_it=100):
        """run SSS_amg_solve one V-cycle per call (max_it = 1) to read ||r|| at full precision;
        the solve has no state besides x, so the sequence equals one call with max_it cycles"""
        x = np.array(x0, np.float64, copy=True)
        bb = np.ascontiguousarray(b, np.float64)
        vx, vb = capi.vec_from_array(x), capi.vec_from_array(bb)
        mg.pars.max_it = 1
        mg.pars.tol = tol
        hist = []
        with quiet():
            for _ in range(max_it):
                r = self.L.SSS_amg_solve(C.byref(mg), C.byref(vx), C.byref(vb))
                hist.append(r.ares)
                if r.rres < tol:
                    break
        return x, np.array(hist)

    def solve_timed(self, mg, x0, b, tol, max_it=100):
        """one SSS_amg_solve call, wall-clocked like the reference does (SSS_SOLVE.c:31,82)"""
        import time
        x = np.array(x0, np.float64, copy=True)
        bb = np.ascontiguousarray(b, np.float64)
        vx, vb = capi.vec_from_array(x), capi.vec_from_array(bb)
        mg.pars.max_it = max_it
        mg.pars.tol = tol
        with quiet():
            t0 = time.perf_counter()
            r = self.L.SSS_amg_solve(C.byref(mg), C.byref(vx), C.byref(vb))
            t1 = time.perf_counter()
        return r, x, t1 - t0
