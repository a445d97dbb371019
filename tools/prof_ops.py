#!/usr/bin/env python3
"""ncu target: the level ops (0 GS sweep, 1 residual, 2 restrict, 3 prolong-add, 4 y = A x, 6 fused residual + restriction) of one level, in
this order, each launched 3 times (2 warm-up + 1 timed by amgb200_time_op).   python tools/prof_ops.py p3d 256 0 [ops]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from amg_b200 import DeviceHierarchy, HostHierarchy, generate
kind, N, level = sys.argv[1], int(sys.argv[2]), int(sys.argv[3])
ops = [int(t) for t in sys.argv[4].split(",")] if len(sys.argv) > 4 else [0, 1, 2, 3, 4, 6]
eps = float(sys.argv[5]) if len(sys.argv) > 5 else 1e-3
hier = HostHierarchy(generate(kind, N, eps), tol=1e-8)
dev = DeviceHierarchy(hier)
names = ["gs_sweep", "residual", "restrict", "prolong", "spmv", "", "resid_restrict"]
for op in ops:
    ms = dev.time_op(level, op, 1)
    print(f"{kind}{N} level {level} {names[op]} ({dev.gs_kernel(level) if op == 0 else 'resid_restrict_kernel' if op == 6 and dev.fused(level) else 'spmv_kernel'}): {ms*1e3:.1f} us, {dev.bytes(level, op)/1e6:.1f} MB algorithmic, {dev.bytes(level, op)/ms/1e6:.0f} GB/s", flush=True)
