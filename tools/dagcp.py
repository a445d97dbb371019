#!/usr/bin/env python3
"""Critical-path analysis of the ordered Gauss-Seidel sweeps (CPU only, no GPU needed).

For every smoothed level of a workload this compares, in units of fp64 chain terms and cycles,
  wf_sum    : sum over wavefronts of the longest row suffix (what the wavefront-barrier kernels pay)
  dag(H)    : the longest path of the fine-grained dependency DAG when every row is an in-order chain
              (c cycles per term) and a finished x becomes usable by a dependent term H cycles after
              the chain of its row ended (divide + store + flag + load + multiply)
usage: python tools/dagcp.py p3d 128
"""
import ctypes as C
import os
import subprocess
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

SRC = r"""
#include <stdlib.h>
#include <string.h>
static double dmaxd(double a, double b) { return a > b ? a : b; }
/* one GS-CF sweep (F pass then C pass); returns finish time (cycles) of the sweep's last row.
   ready[] holds the time x_j becomes usable. c = cycles per chain term, H = hop latency.
   rows of a pass in ascending index; dependency on j in the same pass with j < i; anti-dependencies ignored
   (they only constrain writes, which a dataflow kernel honours by double-buffering or by read-before-write flags). */
double dag_sweep(int n, const int *rp, const int *ci, const int *mark, double c, double H, double *ready, long long *terms_on_path,
                 double *wf_sum_terms, int *nwf) {
    double tend = 0.0;
    int *wf = (int *)calloc(n, sizeof(int));
    int *wfmax = (int *)calloc(2 * n + 2, sizeof(int));
    double base = 0.0;
    int W = 0;
    double wsum = 0;
    for (int pass = 0; pass < 2; ++pass) {
        int wbase = W;
        double pend = base;
        for (int i = 0; i < n; ++i) {
            if ((mark[i] == 1) != pass) continue;
            double t = base;
            int w = 0, firstlate = -1;
            for (int k = rp[i]; k < rp[i + 1]; ++k) {
                const int j = ci[k];
                if (j == i) continue;
                if (((mark[j] == 1) == pass) && j < i) { t = dmaxd(t, ready[j]); if (wf[j] + 1 > w) w = wf[j] + 1; }
                t += c;
            }
            wf[i] = w;
            /* suffix length: entries from the first one whose column is in wavefront w-1 */
            int suf = 0;
            if (w > 0) {
                int len = 0;
                for (int k = rp[i]; k < rp[i + 1]; ++k) {
                    const int j = ci[k];
                    if (j == i) continue;
                    if (firstlate < 0 && ((mark[j] == 1) == pass) && j < i && wf[j] == w - 1) firstlate = len;
                    ++len;
                }
                suf = len - firstlate;
            } else suf = 0;
            if (suf > wfmax[wbase + w]) wfmax[wbase + w] = suf;
            if (wbase + w + 1 > W) W = wbase + w + 1;
            ready[i] = t + H;
            if (t > pend) pend = t;
        }
        base = pend + H;   /* pass boundary: everything visible */
        tend = pend;
    }
    for (int w = 0; w < W; ++w) wsum += wfmax[w];
    *wf_sum_terms = wsum;
    *nwf = W;
    free(wf); free(wfmax);
    (void)terms_on_path;
    return tend;
}
"""


def helper():
    d = tempfile.mkdtemp()
    open(os.path.join(d, "dag.c"), "w").write(SRC)
    so = os.path.join(d, "dag.so")
    subprocess.check_call(["gcc", "-O2", "-shared", "-fPIC", "-o", so, os.path.join(d, "dag.c")])
    L = C.CDLL(so)
    L.dag_sweep.restype = C.c_double
    L.dag_sweep.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_double, C.c_double, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    return L


def main():
    from amg_b200 import HostHierarchy, generate
    kind, N = sys.argv[1], int(sys.argv[2])
    A = generate(kind, N)
    hier = HostHierarchy(A, tol=1e-8)
    L = helper()
    c = 8.6
    ghz = 1.92
    print(f"{'lvl':>3} {'rows':>9} {'nnz':>10} {'wf':>6} {'wf_sum_terms':>12} {'wf_ms(c)':>9} {'wf_ms(+600)':>11} " + " ".join(f"dag_ms(H={h})" for h in (0, 100, 200, 400, 800)))
    tot = np.zeros(7)
    for l in range(hier.num_levels - 1):
        M = hier.level_matrix(l)
        mark = np.ascontiguousarray(hier.cfmark(l), np.int32)
        n = M.nrows
        out = []
        wsum = C.c_double(0); nwf = C.c_int(0)
        for H in (0, 100, 200, 400, 800):
            ready = np.zeros(n)
            t = L.dag_sweep(n, M.row_ptr.ctypes.data, M.col_idx.ctypes.data, mark.ctypes.data, c, float(H), ready.ctypes.data, None, C.byref(wsum), C.byref(nwf))
            out.append(t / ghz / 1e6)
        wf_ms = wsum.value * c / ghz / 1e6
        wf_ms2 = (wsum.value * c + 600.0 * nwf.value) / ghz / 1e6
        tot += np.array([wf_ms, wf_ms2] + out)
        print(f"{l:>3} {n:>9} {M.nnz:>10} {nwf.value:>6} {int(wsum.value):>12} {wf_ms:>9.3f} {wf_ms2:>11.3f} " + " ".join(f"{v:>12.3f}" for v in out))
    print("sum per sweep (ms):", " ".join(f"{v:.3f}" for v in tot), " -> x4 sweeps per V-cycle")


if __name__ == "__main__":
    main()
