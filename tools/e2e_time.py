#!/usr/bin/env python3
"""Developer probe: breakdown of the end-to-end SSS_amg_solve call (host hierarchy in, solution out)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
if len(sys.argv) > 3 and sys.argv[3] == "torch":
    import torch
    torch.cuda.set_device(0); torch.ones(4, device="cuda")
from amg_b200 import HostHierarchy, generate, solve_dropin
kind, N = sys.argv[1], int(sys.argv[2])
A = generate(kind, N); hier = HostHierarchy(A, tol=1e-8); n = A.nrows
os.environ["AMGB200_VERBOSE"] = os.environ.get("AMGB200_VERBOSE", "2")
for rep in range(int(os.environ.get("E2E_REPS", "3"))):
    t = time.perf_counter(); rtn, x = solve_dropin(hier, np.ones(n), np.ones(n)); dt = time.perf_counter() - t
    print(f"=== e2e call {rep}: {dt*1e3:.1f} ms, {rtn.nits} cycles", flush=True)
