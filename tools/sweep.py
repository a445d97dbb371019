#!/usr/bin/env python3
"""Developer probe: ms per smoother sweep of the given levels under several env-knob settings (one host setup).
   python tools/sweep.py p3d 128 4,5,6 "AMGB200_STREAM_S=16" "AMGB200_STREAM_S=8 AMGB200_STREAM_G=2" ..."""
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
    res = []
    for l in levels:
        dev.time_op(l, 0, 2)
        res.append(f"L{l} {dev.time_op(l, 0, 5):.3f}")
    print(f"[{os.environ.get('AMGB200_LIB', 'default').split('/')[-1]}] {st or 'defaults':40s} ms/sweep: " + "  ".join(res), flush=True)
    dev.close()
    for k in kv:
        del os.environ[k]
