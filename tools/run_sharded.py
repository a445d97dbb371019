#!/usr/bin/env python3
"""torchrun target: level-0 sharded solve on N GPUs, checked against the single-GPU solve (bit-identical)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist
from amg_b200 import DeviceHierarchy, HostHierarchy, generate
from amg_b200.distributed import GpuBackend, ShardedSolver

kind = sys.argv[1] if len(sys.argv) > 1 else "p3d"
N = int(sys.argv[2]) if len(sys.argv) > 2 else 32
rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
A = generate(kind, N)
hier = HostHierarchy(A, tol=1e-8)
n = A.nrows
ref = None
if rank == 0:
    dev1 = DeviceHierarchy(hier, device=lr)
    rtn1, x1, hist1 = dev1.solve(np.ones(n), np.ones(n))
    dev1.close()
    ref = (rtn1.nits, x1, hist1)
dev = DeviceHierarchy(hier, device=lr, level0_worker=rank > 0)
solver = ShardedSolver(GpuBackend(dev, torch), A, dist, rank, world)
for rep in range(3):
    dist.barrier(); torch.cuda.synchronize(); t0 = time.perf_counter()
    nits, hist, x = solver.solve(np.ones(n), np.ones(n), 1e-8)
    torch.cuda.synchronize(); dt = time.perf_counter() - t0
    if rank == 0:
        same = x.tobytes() == ref[1].tobytes()
        dev_h = float(np.max(np.abs(hist - ref[2]) / ref[2]))
        print(f"[{world} GPUs] {kind}{N}: {nits} V-cycles (single GPU {ref[0]}), x bit-identical to single GPU: {same}, "
              f"history deviation {dev_h:.1e}, {dt*1e3:.1f} ms, halo {solver.halo_bytes} B/exchange", flush=True)
        assert same and nits == ref[0] and dev_h < 1e-12
dist.destroy_process_group()
