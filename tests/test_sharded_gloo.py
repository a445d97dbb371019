"""Host-side logic of the multi-GPU path (amg_b200/distributed.py) on CPU: world_size-2 gloo processes
run the level-0 sharded V-cycle driver -- partitioning, ghost lists, halo exchange schedule, residual
gather, coarse-correction broadcast -- over a CPU emulation of the device building blocks that is built
on the oracle's row arithmetic.  The sharded solve must reproduce the sequential solve bit for bit."""
import ctypes as C
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


class CpuBackend:
    """the backend interface of amg_b200.distributed.ShardedSolver on host arrays (schedule numbering)"""

    def __init__(self, hier, torch):
        import oracle_ffi
        from amg_b200 import capi
        self.capi, self.torch, self.hier = capi, torch, hier
        self.device = torch.device("cpu")
        self.O = oracle_ffi.Oracle().L
        dp, ip = capi.c_double_p, capi.c_int_p
        self.O.orc_gs_rowlist.argtypes = [C.POINTER(capi.Mat), dp, dp, ip, C.c_int]
        self.O.orc_resid_rowlist.argtypes = [C.POINTER(capi.Mat), dp, dp, dp, ip, C.c_int]
        self.O.orc_amxpy_rowlist.argtypes = [C.c_double, C.POINTER(capi.Mat), dp, dp, ip, C.c_int]
        L = oracle_ffi.test_hooks()
        L.amgb200_debug_schedule.restype = C.c_int
        L.amgb200_debug_schedule.argtypes = [C.POINTER(capi.Mat), ip, ip, ip, C.c_int, ip]
        c0 = hier.level(0)
        n = c0.A.num_rows
        mark = np.ascontiguousarray(hier.cfmark(0))
        order = np.zeros(n, np.int32); wf = np.zeros(n + 2, np.int32); cnt = np.zeros(4, np.int32)
        W = L.amgb200_debug_schedule(C.byref(c0.A), capi.iptr(mark), capi.iptr(order), capi.iptr(wf), n + 2, capi.iptr(cnt))
        self._order = order.astype(np.int64)
        self.n, self.nF = n, int(wf[cnt[0]])
        self.two_colour = (cnt[0], cnt[1]) == (1, 1)
        self.xs = torch.zeros(n, dtype=torch.float64)
        self.bs = torch.zeros(n, dtype=torch.float64)
        self.wps = torch.zeros(n, dtype=torch.float64)
        self.n1 = hier.level(1).A.num_rows
        self.x1t = torch.zeros(self.n1, dtype=torch.float64)

    def shape(self):
        return {"n": self.n, "nF": self.nF, "itemsF": -(-self.nF // 32), "itemsC": -(-(self.n - self.nF) // 32),
                "rows_per_item": 32, "p_items": -(-self.n // 32), "shardable": self.two_colour}

    def order(self): return self._order
    def x0(self): return self.xs
    def wp0(self): return self.wps
    def x1(self): return self.x1t

    def _nat(self, t):
        a = np.empty(self.n)
        a[self._order] = t.numpy()
        return a

    def _rows(self, base, count, a, b):
        k = np.arange(base + 32 * a, base + min(32 * b, count))
        return k, np.ascontiguousarray(self._order[k], np.int32)

    def gs_pass(self, which, a, b):
        base, count = (0, self.nF) if which == 0 else (self.nF, self.n - self.nF)
        k, nat = self._rows(base, count, a, b)
        xn, bn = self._nat(self.xs), self._nat(self.bs)
        self.O.orc_gs_rowlist(C.byref(self.hier.level(0).A), self.capi.dptr(bn), self.capi.dptr(xn), self.capi.iptr(nat), len(nat))
        self.xs[k] = self.torch.from_numpy(xn[nat])

    def residual(self, a, b):
        itemsF = -(-self.nF // 32)
        if a >= itemsF:
            k, nat = self._rows(self.nF, self.n - self.nF, a - itemsF, b - itemsF)
        else:
            k, nat = self._rows(0, self.nF, a, b)
        xn, bn, rn = self._nat(self.xs), self._nat(self.bs), np.zeros(self.n)
        self.O.orc_resid_rowlist(C.byref(self.hier.level(0).A), self.capi.dptr(xn), self.capi.dptr(bn), self.capi.dptr(rn), self.capi.iptr(nat), len(nat))
        self.wps[k] = self.torch.from_numpy(rn[nat])

    def prolong(self, a, b):
        k, nat = self._rows(0, self.n, a, b)
        xn = self._nat(self.xs)
        x1 = np.ascontiguousarray(self.x1t.numpy())
        self.O.orc_amxpy_rowlist(1.0, C.byref(self.hier.level(0).P), self.capi.dptr(x1), self.capi.dptr(xn), self.capi.iptr(nat), len(nat))
        self.xs[k] = self.torch.from_numpy(xn[nat])

    def restrict_and_lower_levels(self):
        capi, mg = self.capi, self.hier.mg
        rn = self._nat(self.wps)
        b1 = np.ctypeslib.as_array(mg.cg[1].b.d, shape=(self.n1,))
        x1 = np.ctypeslib.as_array(mg.cg[1].x.d, shape=(self.n1,))
        self.O.orc_mv_mxy(C.byref(mg.cg[0].R), capi.dptr(rn), mg.cg[1].b.d)
        x1[:] = 0.0
        sub = capi.Amg(mg.num_levels - 1, C.cast(C.addressof(mg.cg[1]), C.POINTER(capi.Comp)), mg.pars, mg.rtn)
        import oracle_ffi
        with oracle_ffi.quiet():
            self.O.orc_cycle(C.byref(sub), 0)
        self.x1t.copy_(self.torch.from_numpy(x1.copy()))
        del b1

    def set_problem(self, x_nat, b_nat):
        self.xs.copy_(self.torch.from_numpy(np.asarray(x_nat, dtype=np.float64)[self._order]))
        self.bs.copy_(self.torch.from_numpy(np.asarray(b_nat, dtype=np.float64)[self._order]))

    def get_solution(self):
        return self._nat(self.xs)


def _worker(rank, world, port, kind, N, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    import torch
    import torch.distributed as dist
    from amg_b200 import HostHierarchy, generate
    from amg_b200.distributed import ShardedSolver
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        A = generate(kind, N)
        hier = HostHierarchy(A, tol=1e-8)
        be = CpuBackend(hier, torch)
        solver = ShardedSolver(be, A, dist, rank, world)
        nits, hist, x = solver.solve(np.ones(A.nrows), np.ones(A.nrows), 1e-8)
        if rank == 0:
            import oracle_ffi
            hier2 = HostHierarchy(A, tol=1e-8)
            rtn, x_o, hist_o = oracle_ffi.Oracle().solve(hier2, np.ones(A.nrows), np.ones(A.nrows), 0)
            q.put((nits, rtn.nits, bool(x.tobytes() == x_o.tobytes()), float(np.max(np.abs(hist - hist_o) / hist_o)), solver.halo_bytes))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("kind,N,world", [("p3d", 12, 2), ("p2d", 50, 2), ("p3d", 16, 3), ("p3d", 10, 1)])
def test_sharded_level0_solve_is_bit_identical_to_sequential(kind, N, world):
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() + N + world) % 2000
    procs = [ctx.Process(target=_worker, args=(r, world, port, kind, N, q)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=300)
        assert p.exitcode == 0
    nits, nits_o, same_x, hist_dev, halo = q.get(timeout=10)
    assert nits == nits_o
    assert same_x, "sharded solution differs from the sequential one"
    assert hist_dev <= 1e-13
    assert halo > 0 or world == 1


def test_partition_and_ghost_lists_cover_every_read():
    """every off-rank column read by a rank's rows appears in exactly one of its ghost lists"""
    from amg_b200 import HostHierarchy, capi, generate
    from amg_b200.distributed import Partition, ghost_lists
    A = generate("p3d", 10)
    hier = HostHierarchy(A, tol=1e-8)
    import torch
    be = CpuBackend(hier, torch)
    sh = be.shape()
    world = 3
    part = Partition(sh["n"], sh["nF"], sh["itemsF"], sh["itemsC"], 32, world)
    order = be.order()
    pos = np.empty(sh["n"], np.int64); pos[order] = np.arange(sh["n"])
    owned_total = 0
    for r in range(world):
        own = np.concatenate([part.rows(r, 0), part.rows(r, 1)])
        owned_total += len(own)
        g = ghost_lists(A, order, part, r)
        ghosts = np.concatenate([v for w in (0, 1) for v in g[w].values()]) if any(g[w] for w in (0, 1)) else np.zeros(0, np.int64)
        assert len(np.unique(ghosts)) == len(ghosts)
        assert not np.intersect1d(ghosts, own).size
        reads = set()
        for k in own:
            i = order[k]
            reads.update(pos[A.col_idx[A.row_ptr[i]:A.row_ptr[i + 1]]].tolist())
        assert reads - set(own.tolist()) == set(ghosts.tolist())
        for w in (0, 1):
            for src, idx in g[w].items():
                assert (part.owner(idx) == src).all()
    assert owned_total == sh["n"]


@pytest.mark.parametrize("kind,N,world", [("p3d", 12, 2), ("p3d", 14, 3), ("p2d", 40, 4), ("p3d", 16, 8)])
def test_native_ghost_lists_equal_the_numpy_ones(kind, N, world):
    """amgb200_ghost_lists (one OpenMP pass in libamgb200.so's host part) against the per-rank and the vectorised numpy versions"""
    from amg_b200 import HostHierarchy, generate
    from amg_b200.distributed import Partition, all_ghost_lists, all_ghost_lists_native, ghost_lists
    A = generate(kind, N)
    hier = HostHierarchy(A, tol=1e-8)
    import torch
    be = CpuBackend(hier, torch)
    sh = be.shape()
    part = Partition(sh["n"], sh["nF"], sh["itemsF"], sh["itemsC"], 32, world)
    order = be.order()
    a, b = all_ghost_lists(A, order, part), all_ghost_lists_native(A, order, part)
    for r in range(world):
        g = ghost_lists(A, order, part, r)
        for w in (0, 1):
            assert sorted(a[r][w]) == sorted(b[r][w]) == sorted(g[w])
            for src in a[r][w]:
                assert np.array_equal(a[r][w][src], b[r][w][src]) and np.array_equal(g[w][src], b[r][w][src])


@pytest.mark.parametrize("kind,N,world", [("p3d", 16, 2), ("p3d", 20, 4), ("p2d", 64, 3)])
def test_interior_items_read_no_ghost(kind, N, world):
    """the interior / boundary split that lets the interior rows of a level-0 pass run while the halo exchange is in flight: no interior item
    holds a row that reads an entry of another rank; interior + boundary cover every item of the rank exactly once"""
    from amg_b200 import HostHierarchy, generate
    from amg_b200.distributed import Partition, all_ghost_lists_native, ghost_lists, interior_split
    A = generate(kind, N)
    hier = HostHierarchy(A, tol=1e-8)
    import torch
    be = CpuBackend(hier, torch)
    sh = be.shape()
    rpi = 32
    part = Partition(sh["n"], sh["nF"], sh["itemsF"], sh["itemsC"], rpi, world)
    order = be.order()
    rg = np.zeros(sh["n"], np.uint8)
    all_ghost_lists_native(A, order, part, rg)
    pos = np.empty(sh["n"], np.int64); pos[order] = np.arange(sh["n"])
    owner = part.owner(np.arange(sh["n"]))
    # reads_ghost against a direct evaluation
    want = np.zeros(sh["n"], np.uint8)
    for i in range(sh["n"]):
        cols = pos[A.col_idx[A.row_ptr[i]:A.row_ptr[i + 1]]]
        want[pos[i]] = (owner[cols] != owner[pos[i]]).any()
    assert np.array_equal(rg, want)
    nF, n = sh["nF"], sh["n"]
    for which, rows0, items in ((0, 0, part.f_items), (1, nF, part.c_items)):
        seg = rg[:nF] if which == 0 else rg[nF:n]
        flags = np.pad(seg, (0, (-len(seg)) % rpi)).reshape(-1, rpi).any(axis=1)
        for r in range(world):
            a, b = items[r]
            (i0, i1), boundary = interior_split(flags, a, b)
            assert not flags[i0:i1].any()
            covered = sorted([(i0, i1)] + boundary)
            covered = [c for c in covered if c[1] > c[0]]
            assert (covered[0][0] == a and covered[-1][1] == b and all(x[1] == y[0] for x, y in zip(covered, covered[1:]))) if b > a else True
            if world > 1 and b - a > 8:
                assert i1 - i0 > 0, "a slab partition must have interior items"
