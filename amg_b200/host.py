"""Python host-side mirror of the reference's call sequence for the solve path.

  A   = generate("p3d", 128)          # SURVEY.md Appendix B operators (C++ generator)
  hier = HostHierarchy(A, tol=1e-8)   # == SSS_amg_setup (from-scratch restatement, bit-identical)
  rtn, x = solve_dropin(hier, x0, b)  # == SSS_amg_solve(mg, x, b): upload + device solve + download
  dev = DeviceHierarchy(hier)         # resident hierarchy: upload once, solve many times

Everything numerical happens inside libamgb200.so; this file only marshals buffers.
"""
import ctypes as C

import numpy as np

from . import capi

KINDS = {"p2d": 0, "p3d": 1, "aniso3d": 2, "v27": 3}


class CsrMatrix:
    """Host CSR matrix owning its arrays (int32 indices, float64 values)."""

    def __init__(self, row_ptr, col_idx, val, ncols=None):
        self.row_ptr = np.ascontiguousarray(row_ptr, np.int32)
        self.col_idx = np.ascontiguousarray(col_idx, np.int32)
        self.val = np.ascontiguousarray(val, np.float64)
        self.nrows = len(self.row_ptr) - 1
        self.ncols = self.nrows if ncols is None else ncols
        self.c = capi.Mat(self.nrows, self.ncols, len(self.col_idx), capi.iptr(self.row_ptr),
                          capi.iptr(self.col_idx), capi.dptr(self.val))

    @property
    def nnz(self):
        return len(self.col_idx)

    def matvec(self, x):
        """numpy reference product (not bit-ordered) for property checks"""
        prod = self.val * x[self.col_idx]
        return np.add.reduceat(prod, self.row_ptr[:-1]) if self.nnz else np.zeros(self.nrows)


def generate(kind, N, eps_z=1e-3):
    """Synthetic level-0 operator (amgb200_generate); returns CsrMatrix."""
    L = capi.lib()
    m = capi.Mat()
    rc = L.amgb200_generate(KINDS[kind], int(N), float(eps_z), C.byref(m))
    if rc != 0:
        raise ValueError(f"amgb200_generate({kind},{N}) failed: {rc}")
    rp, ci, va = capi.mat_arrays(m)
    out = CsrMatrix(rp.copy(), ci.copy(), va.copy())
    L.amgb200_mat_free(C.byref(m))
    return out


def read_mtx_fast(path):
    """amgb200_read_mtx: the C loader fast path (one multi-threaded pass, the reference loader's semantics); returns CsrMatrix"""
    L = capi.lib()
    m = capi.Mat()
    rc = L.amgb200_read_mtx(str(path).encode(), C.byref(m))
    if rc != 0:
        raise ValueError(f"amgb200_read_mtx({path}) failed: {rc}")
    rp, ci, va = capi.mat_arrays(m)
    out = CsrMatrix(rp.copy(), ci.copy(), va.copy(), ncols=m.num_cols)
    L.amgb200_mat_free(C.byref(m))
    return out


def read_mtx(path):
    """MatrixMarket coordinate reader with the reference loader's semantics
    (amg/mmio_highlevel.h:144-305): entries stay in file order inside each row, symmetric
    files are expanded entry by entry, pattern -> 1.0, no sorting, no duplicate merging."""
    with open(path) as f:
        banner = f.readline().lower().split()
        sym = banner[4] in ("symmetric", "hermitian")
        field = banner[3]
        line = f.readline()
        while line.startswith("%"):
            line = f.readline()
        m, n, nz = (int(t) for t in line.split())
        data = np.loadtxt(f, ndmin=2)
    ri = data[:, 0].astype(np.int64) - 1
    cj = data[:, 1].astype(np.int64) - 1
    va = np.ones(nz) if field == "pattern" else data[:, 2].astype(np.float64)
    rows, cols, vals = [], [], []
    if sym:
        off = ri != cj
        seq_r = np.empty(nz + off.sum(), np.int64); seq_c = np.empty_like(seq_r); seq_v = np.empty(len(seq_r))
        # file order, each off-diagonal entry immediately followed by its mirror
        idx = np.arange(nz) + np.concatenate(([0], np.cumsum(off)[:-1]))
        seq_r[idx] = ri; seq_c[idx] = cj; seq_v[idx] = va
        midx = idx[off] + 1
        seq_r[midx] = cj[off]; seq_c[midx] = ri[off]; seq_v[midx] = va[off]
        rows, cols, vals = seq_r, seq_c, seq_v
    else:
        rows, cols, vals = ri, cj, va
    order = np.argsort(rows, kind="stable")
    rp = np.zeros(m + 1, np.int64)
    np.add.at(rp, rows + 1, 1)
    rp = np.cumsum(rp)
    return CsrMatrix(rp, cols[order], vals[order], ncols=n)


class HostHierarchy:
    """Host AMG hierarchy (an SSS_AMG) built by amgb200_setup."""

    def __init__(self, A, tol=1e-8, verbose=0, device_interp=False, device_rap=False, **par_overrides):
        """device_interp: interpolation weights + truncation on the GPU (amgb200_interp_device; same hierarchy bit for bit)
        device_rap: R = P^T and the Galerkin product on the GPU (amgb200_rap_device; same arrays entry for entry)"""
        self.A = A
        self.pars = capi.default_pars(tol)
        for k, v in par_overrides.items():
            setattr(self.pars, k, v)
        self.mg = capi.Amg()
        self._lib = capi.lib()
        if device_interp or device_rap:
            self._lib.amgb200_setup_ex(C.byref(self.mg), C.byref(A.c), C.byref(self.pars), int(verbose), (1 if device_interp else 0) | (2 if device_rap else 0))
        else:
            self._lib.amgb200_setup(C.byref(self.mg), C.byref(A.c), C.byref(self.pars), int(verbose))
        self._alive = True

    @property
    def num_levels(self):
        return self.mg.num_levels

    def level(self, l):
        return self.mg.cg[l]

    def level_matrix(self, l, which="A"):
        m = getattr(self.mg.cg[l], which)
        rp, ci, va = capi.mat_arrays(m)
        return CsrMatrix(rp, ci, va, ncols=m.num_cols)

    def cfmark(self, l):
        c = self.mg.cg[l]
        return np.ctypeslib.as_array(c.cfmark.d, shape=(c.A.num_rows,))

    def table(self):
        return [(self.mg.cg[l].A.num_rows, self.mg.cg[l].A.num_nnzs) for l in range(self.num_levels)]

    def close(self):
        if self._alive:
            self._lib.amgb200_amg_destroy(C.byref(self.mg))
            self._alive = False

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def solve_dropin(hier, x0, b):
    """The reference-facing call: SSS_amg_solve(mg, x, b) with host buffers."""
    L = capi.lib()
    x = np.array(x0, np.float64, copy=True)
    bb = np.ascontiguousarray(b, np.float64)
    vx, vb = capi.vec_from_array(x), capi.vec_from_array(bb)
    rtn = L.SSS_amg_solve(C.byref(hier.mg), C.byref(vx), C.byref(vb))
    # level 0 of the hierarchy now aliases x/b (reference behaviour); detach before they die
    hier.mg.cg[0].x = capi.Vec(0, None)
    hier.mg.cg[0].b = capi.Vec(0, None)
    return rtn, x


class DeviceHierarchy:
    """Resident device mirror (amgb200_upload)."""

    def __init__(self, hier, coarse_mode=0, verbose=0, device=-1, fast=None, level0_worker=False):
        """level0_worker: only what the row-block kernels of level 0 need becomes resident (ranks >= 1 of the sharded solve)"""
        self._lib = capi.lib()
        opt = capi.Options()
        self._lib.amgb200_default_options(C.byref(opt))
        opt.coarse_mode = coarse_mode
        opt.level0_worker = 1 if level0_worker else 0
        if fast is not None:
            opt.fast = int(fast)
        opt.verbose = verbose
        opt.device = device
        self.hier = hier
        self.h = self._lib.amgb200_upload(C.byref(hier.mg), C.byref(opt))
        self.num_levels = self._lib.amgb200_num_levels(self.h)

    def close(self):
        if self.h:
            self._lib.amgb200_free(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def info(self, l):
        a = (C.c_longlong * 8)()
        self._lib.amgb200_level_info(self.h, l, a)
        keys = ("rows", "nnz", "wf_F", "wf_C", "kind", "rows_F", "P_nnz", "R_nnz")
        return dict(zip(keys, list(a)))

    def solve(self, x0, b, hist_cap=200):
        x = np.array(x0, np.float64, copy=True)
        bb = np.ascontiguousarray(b, np.float64)
        hist = np.zeros(hist_cap)
        rtn = self._lib.amgb200_solve(self.h, capi.dptr(x), capi.dptr(bb), capi.dptr(hist), hist_cap)
        return rtn, x, hist[:rtn.nits].copy()

    def solve_device(self, d_x_ptr, d_b_ptr, hist_cap=200):
        hist = np.zeros(hist_cap)
        rtn = self._lib.amgb200_solve_device(self.h, d_x_ptr, d_b_ptr, capi.dptr(hist), hist_cap)
        return rtn, hist[:rtn.nits].copy()

    def cycle(self, x0, b):
        x = np.array(x0, np.float64, copy=True)
        bb = np.ascontiguousarray(b, np.float64)
        self._lib.amgb200_cycle(self.h, capi.dptr(x), capi.dptr(bb))
        return x

    def spmv(self, l, which, x, y=None, alpha=1.0):
        """which: 'A' | 'P' | 'R'.  y None -> y = M x ; else y += alpha M x (copy returned)."""
        w = {"A": 0, "P": 1, "R": 2}[which]
        info_out = self.info(l + 1)["rows"] if which == "R" else self.info(l)["rows"]
        xx = np.ascontiguousarray(x, np.float64)
        if y is None:
            out = np.zeros(info_out)
            self._lib.amgb200_level_spmv(self.h, l, w, 1.0, capi.dptr(xx), 0, capi.dptr(out))
        else:
            out = np.array(y, np.float64, copy=True)
            self._lib.amgb200_level_spmv(self.h, l, w, float(alpha), capi.dptr(xx), 1, capi.dptr(out))
        return out

    def smooth(self, l, nsweeps, x0, b):
        x = np.array(x0, np.float64, copy=True)
        bb = np.ascontiguousarray(b, np.float64)
        self._lib.amgb200_level_smooth(self.h, l, int(nsweeps), capi.dptr(x), capi.dptr(bb))
        return x

    def residual(self, l, x, b):
        xx = np.ascontiguousarray(x, np.float64)
        bb = np.ascontiguousarray(b, np.float64)
        r = np.zeros(len(bb))
        nrm = self._lib.amgb200_level_residual(self.h, l, capi.dptr(xx), capi.dptr(bb), capi.dptr(r))
        return r, nrm

    def resid_restrict(self, l, x, b):
        """r = b - A_l x, bc = R_l r the way the cycle computes them; returns (r, bc, fused?)"""
        xx = np.ascontiguousarray(x, np.float64)
        bb = np.ascontiguousarray(b, np.float64)
        r = np.zeros(len(bb))
        bc = np.zeros(self.info(l + 1)["rows"])
        fused = self._lib.amgb200_level_resid_restrict(self.h, l, capi.dptr(xx), capi.dptr(bb), capi.dptr(r), capi.dptr(bc))
        return r, bc, bool(fused)

    def coarse_solve(self, x0, b, tol):
        x = np.array(x0, np.float64, copy=True)
        bb = np.ascontiguousarray(b, np.float64)
        its = (C.c_int * 2)()
        st = self._lib.amgb200_coarse_solve(self.h, capi.dptr(x), capi.dptr(bb), float(tol), its)
        return st, x, (its[0], its[1])

    def bytes(self, l, op):
        return self._lib.amgb200_algorithmic_bytes(self.h, l, op)

    def time_op(self, l, op, reps=20):
        return self._lib.amgb200_time_op(self.h, l, op, reps)

    def phase_ms(self):
        a = (C.c_double * 8)()
        self._lib.amgb200_last_phase_ms(self.h, a)
        return list(a)

    def level_ms(self, l):
        a = (C.c_double * 4)()
        self._lib.amgb200_last_level_ms(self.h, l, a)
        return list(a)

    def set_profile(self, on):
        self._lib.amgb200_set_profile(self.h, int(on))

    def upload_seconds(self):
        a = (C.c_double * 2)()
        self._lib.amgb200_upload_seconds(self.h, a)
        return a[0], a[1]

    def device_bytes(self):
        return self._lib.amgb200_device_bytes(self.h)

    def chain_terms(self, l):
        return self._lib.amgb200_level_chain_terms(self.h, l)

    def fused(self, l):
        return bool(self._lib.amgb200_level_fused(self.h, l))

    def gs_kernel(self, l):
        return self._lib.amgb200_level_kernel(self.h, l).decode()

    def bench_solve(self, d_x0_ptr, d_b_ptr, d_x_ptr, warmup, steps):
        ms = C.c_double(0)
        rtn = capi.Rtn()
        self._lib.amgb200_bench_solve(self.h, d_x0_ptr, d_b_ptr, d_x_ptr, warmup, steps, C.byref(ms), C.byref(rtn))
        return ms.value, rtn
