#!/usr/bin/env python3
"""Developer probe: host setup vs setup with the device steps (interpolation, transpose + Galerkin product), per-level times.
   python tools/setup_probe.py p3d 128"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from amg_b200 import HostHierarchy, generate
kind, N = sys.argv[1], int(sys.argv[2])
A = generate(kind, N)
for label, kw in (("host", {}), ("device interp + rap", dict(device_interp=True, device_rap=True)), ("device interp + rap (2nd)", dict(device_interp=True, device_rap=True))):
    t = time.time()
    h = HostHierarchy(A, tol=1e-8, verbose=2, **kw)
    print(f"=== {label}: setup {time.time() - t:.2f} s, {h.num_levels} levels", flush=True)
    h.close()
