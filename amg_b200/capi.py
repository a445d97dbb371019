"""ctypes view of the C ABI declared in include/amg_b200.h.

The struct classes are layout-compatible with the reference's SSS_* types
(amg/SSS_main.h:95-251), so the same classes are used to drive the product
(libamgb200.so) and -- from tests/ and bench.py's CPU-baseline leg only -- the
checker libraries under oracle/.
"""
import ctypes as C
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB_PATH = os.environ.get("AMGB200_LIB") or os.path.join(ROOT, "amg_b200", "libamgb200.so")   # AMGB200_LIB: developer builds (e.g. -DAMGB200_TIMELINE)

c_int_p = C.POINTER(C.c_int)
c_double_p = C.POINTER(C.c_double)


class Mat(C.Structure):          # SSS_MAT
    _fields_ = [("num_rows", C.c_int), ("num_cols", C.c_int), ("num_nnzs", C.c_int),
                ("row_ptr", c_int_p), ("col_idx", c_int_p), ("val", c_double_p)]


class Vec(C.Structure):          # SSS_VEC
    _fields_ = [("n", C.c_int), ("d", c_double_p)]


class IVec(C.Structure):         # SSS_IVEC
    _fields_ = [("n", C.c_int), ("d", c_int_p)]


class Rtn(C.Structure):          # SSS_RTN
    _fields_ = [("ares", C.c_double), ("rres", C.c_double), ("nits", C.c_int)]


class Pars(C.Structure):         # SSS_AMG_PARS
    _fields_ = [("cycle_type", C.c_int), ("tol", C.c_double), ("ctol", C.c_double), ("max_it", C.c_int),
                ("cs_type", C.c_int), ("max_levels", C.c_int), ("coarse_dof", C.c_int),
                ("smoother", C.c_int), ("relax", C.c_double), ("cf_order", C.c_int),
                ("pre_iter", C.c_int), ("post_iter", C.c_int), ("poly_deg", C.c_int),
                ("interp_type", C.c_int), ("strong_threshold", C.c_double),
                ("max_row_sum", C.c_double), ("trunc_threshold", C.c_double)]


class Comp(C.Structure):         # SSS_AMG_COMP
    _fields_ = [("A", Mat), ("R", Mat), ("P", Mat), ("b", Vec), ("x", Vec), ("cfmark", IVec), ("wp", Vec)]


class Amg(C.Structure):          # SSS_AMG
    _fields_ = [("num_levels", C.c_int), ("cg", C.POINTER(Comp)), ("pars", Pars), ("rtn", Rtn)]


class Smtr(C.Structure):         # SSS_SMTR
    _fields_ = [("smoother", C.c_int), ("A", C.POINTER(Mat)), ("b", C.POINTER(Vec)), ("x", C.POINTER(Vec)),
                ("relax", C.c_double), ("nsweeps", C.c_int), ("istart", C.c_int), ("iend", C.c_int),
                ("istep", C.c_int), ("ndeg", C.c_int), ("cf_order", C.c_int), ("ordering", c_int_p)]


class Options(C.Structure):      # amgb200_options
    _fields_ = [("coarse_mode", C.c_int), ("verbose", C.c_int), ("device", C.c_int), ("fast", C.c_int),
                ("level0_worker", C.c_int), ("reserved", C.c_int * 3)]


assert C.sizeof(Mat) == 40 and C.sizeof(Vec) == 16 and C.sizeof(Rtn) == 24
assert C.sizeof(Pars) == 104 and C.sizeof(Comp) == 184 and C.sizeof(Amg) == 144 and C.sizeof(Smtr) == 72

EXPORTS = [
    "SSS_amg_solve", "SSS_amg_cycle", "SSS_amg_coarest_solve", "SSS_amg_smoother_pre", "SSS_amg_smoother_post",
    "amgb200_blas_mv_mxy", "amgb200_blas_mv_amxpy", "amgb200_default_options", "amgb200_upload", "amgb200_free",
    "amgb200_solve", "amgb200_solve_device", "amgb200_cycle", "amgb200_level_spmv", "amgb200_level_smooth",
    "amgb200_level_residual", "amgb200_coarse_solve", "amgb200_num_levels", "amgb200_level_info",
    "amgb200_algorithmic_bytes", "amgb200_time_op", "amgb200_launch_count", "amgb200_last_phase_ms",
    "amgb200_version", "amgb200_generate", "amgb200_mat_free", "amgb200_setup", "amgb200_amg_destroy",
    "amgb200_default_pars", "amgb200_last_level_ms", "amgb200_set_profile", "amgb200_upload_seconds",
    "amgb200_device_bytes", "amgb200_level_kernel", "amgb200_level_chain_terms", "amgb200_bench_solve",
    "amgb200_set_stream", "amgb200_level_vec", "amgb200_level_order", "amgb200_l0_shape", "amgb200_l0_gs_pass",
    "amgb200_l0_residual", "amgb200_l0_prolong", "amgb200_restrict_from", "amgb200_cycle_from",
    "amgb200_vec_to_schedule", "amgb200_vec_to_natural", "amgb200_sync", "amgb200_setup_ex", "amgb200_interp_device",
    "amgb200_ipc_export", "amgb200_ipc_open", "amgb200_peer_plan", "amgb200_peer_run", "amgb200_read_mtx", "amgb200_level_download",
    "amgb200_level_resid_restrict", "amgb200_level_fused", "amgb200_rap_device", "amgb200_ghost_lists", "amgb200_ghost_lists_ex", "amgb200_peer_start", "amgb200_peer_wait",
]

_lib = None


def lib():
    """Load libamgb200.so (fails loudly when it has not been built)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(f"{LIB_PATH} is missing: run `make` (or __graft_entry__.build()); there is no fallback")
        L = C.CDLL(LIB_PATH)
        L.SSS_amg_solve.restype = Rtn
        L.SSS_amg_solve.argtypes = [C.POINTER(Amg), C.POINTER(Vec), C.POINTER(Vec)]
        L.SSS_amg_cycle.argtypes = [C.POINTER(Amg)]
        L.SSS_amg_coarest_solve.argtypes = [C.POINTER(Mat), C.POINTER(Vec), C.POINTER(Vec), C.c_double]
        L.SSS_amg_smoother_pre.argtypes = [C.POINTER(Smtr)]
        L.SSS_amg_smoother_post.argtypes = [C.POINTER(Smtr)]
        L.amgb200_blas_mv_mxy.argtypes = [C.POINTER(Mat), C.POINTER(Vec), C.POINTER(Vec)]
        L.amgb200_blas_mv_amxpy.argtypes = [C.c_double, C.POINTER(Mat), C.POINTER(Vec), C.POINTER(Vec)]
        L.amgb200_default_options.argtypes = [C.POINTER(Options)]
        L.amgb200_upload.restype = C.c_void_p
        L.amgb200_upload.argtypes = [C.POINTER(Amg), C.POINTER(Options)]
        L.amgb200_free.argtypes = [C.c_void_p]
        L.amgb200_solve.restype = Rtn
        L.amgb200_solve.argtypes = [C.c_void_p, c_double_p, c_double_p, c_double_p, C.c_int]
        L.amgb200_solve_device.restype = Rtn
        L.amgb200_solve_device.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, c_double_p, C.c_int]
        L.amgb200_cycle.argtypes = [C.c_void_p, c_double_p, c_double_p]
        L.amgb200_level_spmv.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_double, c_double_p, C.c_int, c_double_p]
        L.amgb200_level_smooth.argtypes = [C.c_void_p, C.c_int, C.c_int, c_double_p, c_double_p]
        L.amgb200_level_residual.restype = C.c_double
        L.amgb200_level_residual.argtypes = [C.c_void_p, C.c_int, c_double_p, c_double_p, c_double_p]
        L.amgb200_level_resid_restrict.restype = C.c_int
        L.amgb200_level_resid_restrict.argtypes = [C.c_void_p, C.c_int, c_double_p, c_double_p, c_double_p, c_double_p]
        L.amgb200_coarse_solve.restype = C.c_int
        L.amgb200_coarse_solve.argtypes = [C.c_void_p, c_double_p, c_double_p, C.c_double, c_int_p]
        L.amgb200_num_levels.restype = C.c_int
        L.amgb200_num_levels.argtypes = [C.c_void_p]
        L.amgb200_level_info.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_longlong)]
        L.amgb200_algorithmic_bytes.restype = C.c_double
        L.amgb200_algorithmic_bytes.argtypes = [C.c_void_p, C.c_int, C.c_int]
        L.amgb200_time_op.restype = C.c_double
        L.amgb200_time_op.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int]
        L.amgb200_launch_count.restype = C.c_longlong
        L.amgb200_last_phase_ms.argtypes = [C.c_void_p, c_double_p]
        L.amgb200_version.restype = C.c_char_p
        L.amgb200_generate.restype = C.c_int
        L.amgb200_generate.argtypes = [C.c_int, C.c_int, C.c_double, C.POINTER(Mat)]
        L.amgb200_mat_free.argtypes = [C.POINTER(Mat)]
        L.amgb200_read_mtx.restype = C.c_int
        L.amgb200_read_mtx.argtypes = [C.c_char_p, C.POINTER(Mat)]
        L.amgb200_setup.argtypes = [C.POINTER(Amg), C.POINTER(Mat), C.POINTER(Pars), C.c_int]
        L.amgb200_setup_ex.argtypes = [C.POINTER(Amg), C.POINTER(Mat), C.POINTER(Pars), C.c_int, C.c_int]
        L.amgb200_interp_device.restype = C.c_int
        L.amgb200_interp_device.argtypes = [C.POINTER(Mat), c_int_p, C.POINTER(Mat), C.c_double]
        L.amgb200_amg_destroy.argtypes = [C.POINTER(Amg)]
        L.amgb200_default_pars.argtypes = [C.POINTER(Pars)]
        L.amgb200_last_level_ms.argtypes = [C.c_void_p, C.c_int, c_double_p]
        L.amgb200_set_profile.argtypes = [C.c_void_p, C.c_int]
        L.amgb200_upload_seconds.argtypes = [C.c_void_p, c_double_p]
        L.amgb200_device_bytes.restype = C.c_longlong
        L.amgb200_device_bytes.argtypes = [C.c_void_p]
        L.amgb200_level_chain_terms.restype = C.c_longlong
        L.amgb200_level_chain_terms.argtypes = [C.c_void_p, C.c_int]
        L.amgb200_level_fused.restype = C.c_int
        L.amgb200_level_fused.argtypes = [C.c_void_p, C.c_int]
        L.amgb200_level_kernel.restype = C.c_char_p
        L.amgb200_level_kernel.argtypes = [C.c_void_p, C.c_int]
        L.amgb200_bench_solve.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int,
                                          c_double_p, C.POINTER(Rtn)]
        L.amgb200_set_stream.argtypes = [C.c_void_p, C.c_void_p]
        L.amgb200_level_vec.restype = C.c_void_p
        L.amgb200_level_vec.argtypes = [C.c_void_p, C.c_int, C.c_int]
        L.amgb200_level_order.argtypes = [C.c_void_p, C.c_int, c_int_p]
        L.amgb200_level_download.argtypes = [C.c_void_p, C.c_int, C.c_int, c_double_p]
        L.amgb200_l0_shape.argtypes = [C.c_void_p, C.POINTER(C.c_longlong)]
        L.amgb200_l0_gs_pass.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int]
        L.amgb200_l0_residual.argtypes = [C.c_void_p, C.c_int, C.c_int]
        L.amgb200_l0_prolong.argtypes = [C.c_void_p, C.c_int, C.c_int]
        L.amgb200_restrict_from.argtypes = [C.c_void_p, C.c_int]
        L.amgb200_cycle_from.argtypes = [C.c_void_p, C.c_int]
        L.amgb200_vec_to_schedule.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
        L.amgb200_vec_to_natural.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
        L.amgb200_sync.argtypes = [C.c_void_p]
        L.amgb200_ipc_export.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_char_p]
        L.amgb200_ipc_open.restype = C.c_void_p
        L.amgb200_ipc_open.argtypes = [C.c_void_p, C.c_char_p]
        L.amgb200_peer_plan.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.POINTER(C.c_void_p), C.POINTER(c_int_p), c_int_p,
                                        c_int_p, C.c_int, C.POINTER(C.c_void_p), C.c_int, c_int_p]
        L.amgb200_peer_run.argtypes = [C.c_void_p, C.c_int]
        L.amgb200_peer_start.argtypes = [C.c_void_p, C.c_int]
        L.amgb200_peer_wait.argtypes = [C.c_void_p, C.c_int]
        _lib = L
    return _lib


# ---- numpy <-> struct helpers -------------------------------------------------------------
def dptr(a):
    return a.ctypes.data_as(c_double_p)


def iptr(a):
    return a.ctypes.data_as(c_int_p)


def mat_arrays(m):
    """(row_ptr, col_idx, val) numpy views of a Mat (no copy; valid while the owner lives)."""
    rp = np.ctypeslib.as_array(m.row_ptr, shape=(m.num_rows + 1,))
    nnz = int(rp[-1])
    ci = np.ctypeslib.as_array(m.col_idx, shape=(nnz,)) if nnz else np.zeros(0, np.int32)
    va = np.ctypeslib.as_array(m.val, shape=(nnz,)) if nnz else np.zeros(0)
    return rp, ci, va


def mat_from_arrays(rp, ci, va, ncols):
    """Mat pointing at numpy arrays; returns (Mat, keepalive)."""
    rp = np.ascontiguousarray(rp, np.int32)
    ci = np.ascontiguousarray(ci, np.int32)
    va = np.ascontiguousarray(va, np.float64)
    m = Mat(len(rp) - 1, ncols, len(ci), iptr(rp), iptr(ci), dptr(va))
    return m, (rp, ci, va)


def vec_from_array(a):
    return Vec(len(a), dptr(a))


def default_pars(tol=1e-8):
    p = Pars()
    p.smoother = 2; p.max_it = 100; p.tol = tol; p.ctol = 1e-7; p.max_levels = 30; p.coarse_dof = 10
    p.cycle_type = 1; p.cf_order = 1; p.pre_iter = 2; p.post_iter = 2; p.relax = 1.0; p.poly_deg = 3
    p.cs_type = 1; p.interp_type = 1; p.max_row_sum = 0.9; p.strong_threshold = 0.3; p.trunc_threshold = 0.2
    return p
