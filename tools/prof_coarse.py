#!/usr/bin/env python3
"""ncu target: the coarsest-level solve (one coarse_cg_kernel launch per call) of a workload, 3 calls.   python tools/prof_coarse.py p3d 128"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from amg_b200 import DeviceHierarchy, HostHierarchy, generate
kind, N = sys.argv[1], int(sys.argv[2])
hier = HostHierarchy(generate(kind, N), tol=1e-8)
dev = DeviceHierarchy(hier)
n = hier.level(hier.num_levels - 1).A.num_rows
b = np.random.default_rng(5).standard_normal(n)
for rep in range(3):
    t = time.perf_counter(); st, x, its = dev.coarse_solve(np.zeros(n), b, 1e-9); dt = time.perf_counter() - t
    print(f"coarse solve {rep}: {n} rows, status {st}, its {its}, {dt*1e3:.2f} ms incl. copies", flush=True)
