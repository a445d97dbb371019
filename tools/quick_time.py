#!/usr/bin/env python3
"""Developer timing probe (not the bench): per-level op times and a full solve on one GPU."""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from amg_b200 import DeviceHierarchy, HostHierarchy, generate  # noqa: E402

kind = sys.argv[1] if len(sys.argv) > 1 else "p3d"
N = int(sys.argv[2]) if len(sys.argv) > 2 else 64
eps = float(sys.argv[3]) if len(sys.argv) > 3 else 1e-3
t = time.time(); A = generate(kind, N, eps); print(f"generate {kind} {N}: {time.time()-t:.2f}s n={A.nrows} nnz={A.nnz}", flush=True)
t = time.time(); hier = HostHierarchy(A, tol=1e-8); print(f"setup: {time.time()-t:.2f}s levels={hier.num_levels}", flush=True)
t = time.time(); dev = DeviceHierarchy(hier, verbose=2); print(f"upload+analysis: {time.time()-t:.2f}s", flush=True)
names = ["GS sweep", "residual", "restrict", "prolong", "y=Ax"]
for l in range(dev.num_levels):
    info = dev.info(l)
    row = f"L{l} n={info['rows']:>9} nnz={info['nnz']:>10} wf={info['wf_F']}/{info['wf_C']} kind={info['kind']} {dev.gs_kernel(l)}:"
    for op in range(5):
        ms = dev.time_op(l, op, 5)
        by = dev.bytes(l, op)
        if ms > 0:
            row += f"  {names[op]} {ms*1e3:8.1f}us {by/ms/1e6:7.1f}GB/s"
    print(row, flush=True)
n = A.nrows
dev.set_profile(1)
rtn, x, hist = dev.solve(np.ones(n), np.ones(n))
dev.set_profile(0)
for l in range(dev.num_levels):
    lm = dev.level_ms(l)
    ct = dev.chain_terms(l)
    print(f"  profiled solve L{l}: GS {lm[0]:9.2f} ms  resid {lm[1]:8.2f}  restrict {lm[2]:8.2f}  prolong {lm[3]:8.2f}   chain floor {ct*8.1/1965e3*4*rtn.nits:9.2f} ms", flush=True)
print("  phases [GS,resid,restrict,prolong,coarse,outer,total]:", [round(v, 2) for v in dev.phase_ms()[:7]], "its", rtn.nits, flush=True)
for rep in range(2):
    t = time.time(); rtn, x, hist = dev.solve(np.ones(n), np.ones(n)); dt = time.time() - t
    print(f"solve: {dt*1e3:.1f} ms  its={rtn.nits} relres={rtn.rres:.3e}  {dt*1e3/max(1,rtn.nits):.2f} ms/cycle  vcycle bytes {dev.bytes(0,5)/1e9:.3f} GB", flush=True)
print("phase ms [GS,resid,restrict,prolong,coarse,outer,total]:", [round(v, 2) for v in dev.phase_ms()[:7]])
print("hist", hist)
