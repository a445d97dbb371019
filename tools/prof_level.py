#!/usr/bin/env python3
"""Launch one level's Gauss-Seidel sweep a few times (target for ncu -k regex:gs_)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from amg_b200 import DeviceHierarchy, HostHierarchy, generate
kind, N, level = sys.argv[1], int(sys.argv[2]), int(sys.argv[3])
op = int(sys.argv[4]) if len(sys.argv) > 4 else 0
hier = HostHierarchy(generate(kind, N), tol=1e-8)
dev = DeviceHierarchy(hier, verbose=2)
print("ms per launch:", dev.time_op(level, op, 3))
