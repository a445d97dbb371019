"""Developer probe: scan_fold_slots against the sequential chain (cycles per term), per generator mode and lanes per row."""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from amg_b200 import capi
L = capi.lib()
L.amgb200_debug_scanfold_check.restype = C.c_longlong
L.amgb200_debug_scanfold_check.argtypes = [C.c_int, C.c_int, C.c_ulonglong, C.c_int, C.c_int, C.POINTER(C.c_double)]
cyc = (C.c_double * 2)()
for active in (1,):
    for mode in (0, 5, 2):
        for sub in (32, 8):
            bad = L.amgb200_debug_scanfold_check(148, 64, 7, mode | (active << 8), sub, cyc)
            print(f"active warps/block {active} mode {mode} sub {sub}: bad {bad}  chain {cyc[0]:.2f}  scan {cyc[1]:.2f} cycles/term", flush=True)
