#!/usr/bin/env python3
"""Developer probe: residual, restriction and the fused residual (+) restriction launch of the given levels under several env-knob settings.
   python tools/fused_probe.py p3d 256 0,1 "AMGB200_RR_CHUNKS=128" "AMGB200_RR_LAG=2" "AMGB200_NO_FUSED=1" """
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from amg_b200 import DeviceHierarchy, HostHierarchy, generate
kind, N = sys.argv[1], int(sys.argv[2])
levels = [int(t) for t in sys.argv[3].split(",")]
settings = [""] + sys.argv[4:]
hier = HostHierarchy(generate(kind, N), tol=1e-8)
for st in settings:
    kv = dict(t.split("=") for t in st.split()) if st else {}
    for k, v in kv.items():
        os.environ[k] = v
    dev = DeviceHierarchy(hier)
    for l in levels:
        t1, t2, t6 = (dev.time_op(l, op, 20) for op in (1, 2, 6))
        by = dev.bytes(l, 6)
        print(f"{st or 'defaults':44s} L{l}: residual {t1*1e3:7.1f} us  restrict {t2*1e3:7.1f} us  cycle step (op 6) {t6*1e3:7.1f} us = {by/t6/1e6:7.1f} GB/s of the fused byte count", flush=True)
    dev.close()
    for k in kv:
        del os.environ[k]
