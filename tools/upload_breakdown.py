import sys, time
sys.path.insert(0, ".")
import numpy as np
from amg_b200 import DeviceHierarchy, HostHierarchy, generate
A = generate("p3d", 128)
hier = HostHierarchy(A, tol=1e-8)
for rep in range(3):
    t = time.time(); dev = DeviceHierarchy(hier, verbose=3 if rep == 2 else 0); print("upload", round(time.time() - t, 4), flush=True); dev.close()
