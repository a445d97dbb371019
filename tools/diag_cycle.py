#!/usr/bin/env python3
"""Developer diagnostic: walk V-cycles step by step on GPU (hooks) and oracle, report first deviations."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from amg_b200 import DeviceHierarchy, HostHierarchy, generate
import oracle_ffi
kind = sys.argv[1] if len(sys.argv) > 1 else "p2d"; N = int(sys.argv[2]) if len(sys.argv) > 2 else 256
A = generate(kind, N); hier = HostHierarchy(A, tol=1e-8); dev = DeviceHierarchy(hier); O = oracle_ffi.Oracle()
nl = hier.num_levels; n = A.nrows
def d(a, b):
    if a.tobytes() == b.tobytes(): return "bit-identical"
    return "rel %.2e" % (np.abs(a - b).max() / max(np.abs(b).max(), 1e-300))
xg = [None] * nl; xo = [None] * nl; bg = [None] * nl; bo = [None] * nl
xg[0] = np.ones(n); xo[0] = np.ones(n); bg[0] = np.ones(n); bo[0] = np.ones(n)
tol_c = 1e-9
for cyc in range(1, 10):
    for l in range(nl - 1):
        c = hier.level(l)
        xg[l] = dev.smooth(l, 2, xg[l], bg[l]); xo[l] = O.gs_cf(c.A, hier.cfmark(l), xo[l], bo[l], 2, 1)
        rg, _ = dev.residual(l, xg[l], bg[l]); ro = O.amxpy(-1.0, c.A, xo[l], bo[l])
        bg[l + 1] = dev.spmv(l, "R", rg); bo[l + 1] = O.mxy(c.R, ro)
        xg[l + 1] = np.zeros(len(bg[l + 1])); xo[l + 1] = np.zeros(len(bo[l + 1]))
        print(f"cyc {cyc} down L{l}: x {d(xg[l], xo[l])}  r {d(rg, ro)}  b_next {d(bg[l+1], bo[l+1])}")
    c = hier.level(nl - 1)
    st, xg[nl - 1], its = dev.coarse_solve(xg[nl - 1], bg[nl - 1], tol_c)
    sto, xo[nl - 1], itso = O.coarse_solve(c.A, xo[nl - 1], bo[nl - 1], tol_c, 0)
    print(f"cyc {cyc} coarse: its gpu {its} cpu {itso}  x {d(xg[nl-1], xo[nl-1])}  |b_c| {np.linalg.norm(bo[nl-1]):.3e}")
    for l in range(nl - 2, -1, -1):
        c = hier.level(l)
        xg[l] = dev.spmv(l, "P", xg[l + 1], xg[l], 1.0); xo[l] = O.amxpy(1.0, c.P, xo[l + 1], xo[l])
        pre = d(xg[l], xo[l])
        xg[l] = dev.smooth(l, 2, xg[l], bg[l]); xo[l] = O.gs_cf(c.A, hier.cfmark(l), xo[l], bo[l], 2, 1)
        print(f"cyc {cyc} up   L{l}: after prolong {pre}  after smooth {d(xg[l], xo[l])}")
    rg, ng = dev.residual(0, xg[0], bg[0]); ro = O.amxpy(-1.0, hier.level(0).A, xo[0], bo[0]); no = np.sqrt(np.sum(ro * ro))
    print(f"== cyc {cyc}: |r| gpu {ng:.17e} cpu {no:.17e} rel dev {abs(ng-no)/no:.2e}")
    if no / np.sqrt(n) < 1e-8: break
