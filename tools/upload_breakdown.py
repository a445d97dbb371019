#!/usr/bin/env python3
"""Developer probe: time of amgb200_upload (schedule analysis + layouts + H2D) with its per-step notes.   python tools/upload_breakdown.py p3d 128"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from amg_b200 import DeviceHierarchy, HostHierarchy, generate
kind = sys.argv[1] if len(sys.argv) > 1 else "p3d"
N = int(sys.argv[2]) if len(sys.argv) > 2 else 128
hier = HostHierarchy(generate(kind, N), tol=1e-8)
for rep in range(3):
    t = time.time(); dev = DeviceHierarchy(hier, verbose=3 if rep == 2 else 0); print("upload", round(time.time() - t, 4), flush=True); dev.close()
